"""Times the VCN forward (kernel group) with CUDA events; what-if switches via SEEVCN_CHAIN_DBG."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seevcn_b200 import synth
from seevcn_b200.see.surface_completion.models.vcn.models.build import MODELS
dev = torch.device("cuda", 0)
nobj = int(sys.argv[1]) if len(sys.argv) > 1 else 308
torch.manual_seed(0)
m = MODELS.build({"NAME": "VCN_VC"}, precision="bf16").to(dev).eval()
part, _, _ = synth.make_object_clouds(3, min(nobj, 100), 1024, 0)
x = torch.from_numpy(part).to(dev).repeat((nobj + part.shape[0] - 1) // part.shape[0], 1, 1)[:nobj].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): m({"input": x})
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m({"input": x}); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
ts.sort()
flop = 2.0 * (959040 * 1024 + 5771776) * nobj
print(f"dbg={os.environ.get('SEEVCN_CHAIN_DBG','0')} objs={nobj} forward median {ts[5]*1e3:.1f} us  min {ts[0]*1e3:.1f} us  -> {flop/ts[5]/1e9:.0f} TFLOP/s")
