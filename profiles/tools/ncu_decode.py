"""Three generate() calls at the production size, nothing else (target of the ncu capture of decode_kernel)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, 'lct-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from vqa_model import QstEncoder
torch.manual_seed(0)
q = QstEncoder(17858, 300, 512, 1, 512).to('cuda')
img = torch.randn(64, 512, device='cuda') * 0.1
for _ in range(3):
    out = q.generate(img)
torch.cuda.synchronize()
print(out[0, :8].tolist())
