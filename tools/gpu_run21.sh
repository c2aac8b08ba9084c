#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_memcheck.log 2>&1
( time timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_racecheck.log 2>&1
( time timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_synccheck.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck.log; tail -4 gpurun_out/sanitizer_racecheck.log; tail -4 gpurun_out/sanitizer_synccheck.log
python -m pytest tests/test_parity_at_size.py -m gpu -q -s -k "benchmark_size" 2>&1 | grep -E "parity@size|passed|failed" | tee gpurun_out/parity_at_size_r2.txt
python - <<'PY'
import torch, time
a = torch.empty(47191680 // 4).pin_memory()
torch.cuda.synchronize()
for _ in range(3): b = a.to("cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): b = a.to("cuda")
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"H2D 47 MB pinned .to(cuda): {dt*1e3:.3f} ms = {47.19/dt/1e3:.1f} GB/s")
PY
