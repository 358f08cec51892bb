#!/bin/bash
# Round-end measurement set for one B200 (run through gpurun; everything lands in gpurun_out/).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 400 gpurun_out/bench_final.json; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:^chd_k_kkt$ -s 60 -c 1 -f -o gpurun_out/kkt_final python scripts/gpu_first.py 64 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out/kkt_final.ncu-rep gpurun_out/launches_final.csv
