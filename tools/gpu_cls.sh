#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "cls_attention" 2>&1 | tail -2
python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from sais_b200 import ops
dev = torch.device('cuda:0'); B = 256
qkv = torch.randn(B * 197, 1152, device=dev).bfloat16()
for _ in range(3): ops.vit_cls_attention(qkv, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.vit_cls_attention(qkv, B)
e1.record(); torch.cuda.synchronize()
print("vit_cls_attention", e0.elapsed_time(e1) / 20 * 1e3, "us")
PY
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"
