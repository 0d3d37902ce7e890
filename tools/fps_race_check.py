import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, oracle
from seevcn_b200 import synth, _abi
from seevcn_b200.pcdet.ops.pointnet2.pointnet2_batch import pointnet2_utils as pn2
cuda = torch.device("cuda", 0)
R = oracle.ref_kernels()
for n, m, src in ((16384, 256, 3000), (4096, 128, 300)):
    part, _, _ = synth.make_object_clouds(90 + n, 3, src, 0)
    tiled = np.tile(part, (1, n // src + 1, 1))[:, :n].copy()
    xyz = torch.from_numpy(tiled).to(cuda)
    got = pn2.furthest_point_sample(xyz, m)
    want = oracle.furthest_point_sample(tiled, m)
    print(n, m, "ours == oracle:", np.array_equal(got.cpu().numpy(), want))
    if R is not None:
        ref = torch.empty_like(got)
        temp = torch.full((3, n), 1e10, device=cuda)
        R.ref_fps(3, n, m, _abi.ptr(xyz), _abi.ptr(temp), _abi.ptr(ref))
        print(n, m, "reference kernel == oracle:", np.array_equal(ref.cpu().numpy(), want))
