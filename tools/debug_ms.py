import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nlp_filter_b200 import engine, multiseg, shapes, synth
from oracle import multiseg as oms
d = shapes.vdp_mhe(); B = 24
bt = synth.vdp_mhe_batch(d, B, seed=77)
ms = multiseg.MultiSegmentProblem(shapes.vdp_mhe(), 1)
z = torch.as_tensor(bt["z0"].copy()).cuda()
zz, J, it, st = ms.solve(z.clone(), bt["U"][:, None], bt["data"][:, None])
print("S=1 status", st.cpu().numpy(), "iters", it.cpu().numpy())
seg = shapes.vdp_segment(); S, B = 40, 48
bt = synth.vdp_multiseg_batch(seg, S, B, seed=6000)
ms = multiseg.MultiSegmentProblem(seg, S)
z = torch.as_tensor(bt["z0"].copy()).cuda()
z, J, it, st = ms.solve(z, bt["U"], bt["data"])
print("S=40 status", st.cpu().numpy(), "iters", it.cpu().numpy())
b = int(np.argmax(st.cpu().numpy()))
zo, info = oms.solve(seg, S, bt["z0"][b], bt["U"][b], bt["data"][b])
print("oracle on window", b, info, "zdiff", np.max(np.abs(z.cpu().numpy()[b] - zo)), "J gpu", float(J[b]))
