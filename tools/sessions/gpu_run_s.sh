#!/bin/bash
# session S: parity incl. mstd + ranking, headline bench, ranking timing at E. coli size
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out/s; mkdir -p $O
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "rc=$?"; tail -12 $O/pytest_gpu.log
VARIANTS="base" bash tools/gpu_bench_variants.sh
timeout 600 python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, ".")
import nanomod_b200 as nm
from bench import make_device_workload
det = nm.Detector(0)
L = 4_600_000
dev, _ = make_device_workload(L, 100, 100, torch.device("cuda:0"))
opt = nm.DetectOptions(neighborPvalues=3, testMethod="stouffer", want_u=False, want_t=False)
out = nm.alloc_device_table(opt, L, "cuda:0")
n = det.detect_device(dev, opt, out)
for _ in range(2): order = det.rank_device(out, n, opt)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(5): order = det.rank_device(out, n, opt)
torch.cuda.synchronize(); t_dev = (time.time() - t0) / 5
ks = out["ks_p"].cpu().numpy(); sp = out["stouffer_p"].cpu().numpy()
t0 = time.time(); want = np.lexsort((ks, sp)); t_np = time.time() - t0
print("rank 4.6M rows: device %.2f ms, numpy lexsort %.0f ms, equal %s" % (t_dev * 1e3, t_np * 1e3, np.array_equal(order.cpu().numpy(), want)))
PY
