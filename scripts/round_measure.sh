#!/bin/bash
# Round-end measurement set for one B200 (run through gpurun; everything lands in gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
tail -c 700 gpurun_out/bench_ref.json; echo
timeout 300 python scripts/bench_contact.py > gpurun_out/contact_final.json 2> gpurun_out/contact_final.err
tail -c 500 gpurun_out/contact_final.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^chd_k_kkt -s 40 -c 1 -f -o gpurun_out/kkt_final python scripts/gpu_first.py 64 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/kkt_final.ncu-rep gpurun_out/launches_final.csv
