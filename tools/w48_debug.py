import os, sys
ROOT="/root/repo"
sys.path.insert(0, ROOT+"/arbitrary-hands-3d-reconstruction_b200"); sys.path.insert(0, ROOT)
os.environ["ACR_B200_DEBUG_SYNC"]="1"
import torch
from acr_b200.engine import Engine
from acr_b200.netspec import WIDTHS_W48, build_acr_spec
from acr_b200.synth import synth_state_dict
sd = synth_state_dict(3, spec=build_acr_spec(512, widths=WIDTHS_W48))
gi = torch.Generator().manual_seed(123)
img = torch.randint(0, 256, (1, 512, 512, 3), generator=gi, dtype=torch.uint8).cuda()
eng = Engine(sd, 1, "cuda", torch.bfloat16, widths=WIDTHS_W48)
try:
    eng.run(img); torch.cuda.synchronize(); print("W48 plan ran")
except Exception as e:
    print("FAIL:", e)
