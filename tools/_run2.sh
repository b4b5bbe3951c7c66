timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -8
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
from casualhdrsplat_b200.train import ssim_loss, photometric_loss
dev = torch.device("cuda:0")
x = torch.rand(1, 1080, 1920, 3, device=dev); y = torch.rand(1, 1080, 1920, 3, device=dev)
for name, fn in [("ssim", lambda: ssim_loss(x, y)), ("l2", lambda: photometric_loss(x, y))]:
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "ms per 1080p frame", e0.elapsed_time(e1) / 20)
PY
