MCHRG_B200_LIB=$PWD/multicharge_b200/libmchrg_b200_tmaprof.so MCHRG_B200_TRACE=1 timeout 300 python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import numpy as np, torch
import multicharge_b200 as mc
from synth import batch
ctx = mc.Context(0)
model = mc.new_eeq2019_batch_model(ctx)
offsets, num, xyz, charge = batch(20000, 5, 30, 200)
dev = torch.device("cuda", 0)
d_in = [torch.tensor(a, device=dev) for a in (offsets, num, xyz, charge)]
ntot = int(offsets[-1])
out = dict(qvec=torch.empty(ntot, dtype=torch.float64, device=dev), energy=torch.empty(ntot, dtype=torch.float64, device=dev),
           gradient=torch.empty(ntot, 3, dtype=torch.float64, device=dev), status=torch.empty(20000, dtype=torch.int32, device=dev))
mc.singlepoint_batch(model, *d_in, out=out)
torch.cuda.synchronize()
print(ctx.batch_trace())
PY
