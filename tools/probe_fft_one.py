import os, sys, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from cngi_prototype_b200 import _fft
n_pad, n_img = int(sys.argv[1]), int(sys.argv[2])
g = torch.view_as_complex(torch.randn((1, 2, n_pad, n_pad, 2), dtype=torch.float32, device="cuda"))
for _ in range(2):
    img = _fft.grid_to_image(g, (n_img, n_img))
torch.cuda.synchronize()
