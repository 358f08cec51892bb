"""Contact-net inference benchmark (BASELINE.json configs[2]: 100k 9-frame OpenPose-25 windows, 1 B200).
Prints one JSON line: windows/s of the CUDA path (device-resident and end-to-end with host buffers), the fp32
roofline fraction of the MLP kernel and the torch-CPU reference restatement timed on the host cores."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np
import torch
import chd
from make_contact_golden import contact_weights, synth_keypoints
from oracle import contact as oc

V, F = 1000, 108                      # 1000 videos x 100 windows
rng = np.random.default_rng(0)
base = [synth_keypoints(i, F) for i in range(8)]
raw = [base[i % 8] + rng.normal(0, 0.5, base[0].shape) * np.array([1, 1, 0]) for i in range(V)]
frames, seq_lens = chd.contact.preprocess_videos(raw)
frames = torch.from_numpy(frames).pin_memory().numpy()      # page-locked host buffer, as the bench contract asks for the e2e leg
sd = contact_weights(0)
net = chd.contact.ContactNet(sd)
nwin = V * (F - 8)
for _ in range(3):
    labels, mabs = net.forward(frames, seq_lens)
t0 = time.perf_counter()
K = 5
for _ in range(K):
    labels, mabs = net.forward(frames, seq_lens)
e2e = (time.perf_counter() - t0) / K
# device-resident timing through the *_device entry point
L = net.L
import ctypes as C
L.chd_contact_forward_device.argtypes = [C.c_void_p] * 2 + [C.c_int32] * 2 + [C.c_void_p] * 5
fr = torch.from_numpy(frames).cuda(); sl = torch.from_numpy(seq_lens).cuda()
lab = torch.empty((V, F, 4), dtype=torch.int64, device="cuda"); lg = torch.empty((nwin, 20), dtype=torch.float32, device="cuda")
mn = torch.empty(1, dtype=torch.float32, device="cuda")
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
ts = []
for i in range(8):
    flush.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.chd_contact_forward_device(net.h, fr.data_ptr(), V, F, sl.data_ptr(), lab.data_ptr(), lg.data_ptr(), mn.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    if i >= 3: ts.append(e0.elapsed_time(e1) * 1e-3)
dev = float(np.mean(ts))
assert np.array_equal(lab.cpu().numpy(), labels)
# CPU reference restatement (torch fp32, all host threads) on a bounded sample
ns = 100
t0 = time.perf_counter()
ref_logits = oc.forward_torch(sd, oc.windows_from_frames(frames[:ns]))
ref_lab = [oc.vote(ref_logits[i], int(seq_lens[i])) for i in range(ns)]
cpu = (time.perf_counter() - t0)
agree = np.mean([np.array_equal(ref_lab[i], labels[i, :seq_lens[i]]) for i in range(ns)])
flops = 2 * 953984 * nwin
print(json.dumps({"metric": "contact windows/s", "windows": nwin, "value": nwin / dev, "unit": "windows/s",
                  "e2e": {"value": nwin / e2e, "unit": "windows/s", "h2d_bytes": int(frames.nbytes), "d2h_bytes": int(labels.nbytes)},
                  "roofline": {"bound": "fp32 FFMA (no tensor core: labels must match the fp32 reference)", "achieved_tflops": flops / dev / 1e12,
                               "peak_tflops_nominal": 148 * 128 * 2 * 1.965e9 / 1e12},
                  "cpu_baseline": {"value": ns * (F - 8) / cpu, "unit": "windows/s", "cores": torch.get_num_threads(), "kind": "port",
                                   "sample": "%d videos (%d windows), torch fp32 CPU restatement incl. window building and voting" % (ns, ns * (F - 8))},
                  "labels_equal_frac_vs_cpu": float(agree), "min_abs_logit": mabs, "gpu_launches": int(net.launch_count())}))
