"""Two ViT steps at the bench batch (256 u8 frames) for ncu captures.  Dev tool, GPU only."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import sais_b200.vision_transformer as vits  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vit = vits.vit_small(patch_size=16).to(dev).eval()
fr = torch.randint(0, 256, (B, 224, 224, 3), dtype=torch.uint8, device=dev)
for _ in range(2):
    out = vit.forward_u8(fr)
torch.cuda.synchronize()
print(out.float().abs().mean().item())
