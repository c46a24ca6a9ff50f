import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_py
from common import np_inputs, random_history, run_gpu_resolve
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = 256, 144
sc = SyntheticScene(W, H, pan_px=(5.25, -2.5))
f0, f1 = sc.frame(2), sc.frame(3)
p = abi.TaaParameters.from_buffer_copy(configs.config2_resolve()); p.mRejectOutside = 1
u = configs.uniforms_for(p, f1.jitter_ndc)
ins, hist = np_inputs(f1), random_history(H, W, 11)
want = ("history_out", "result", "mask")
ref = oracle_py.resolve(u, ins["color"], ins["depth"], ins["velocity"], hist, history_depth=f0.depth.numpy(), want=want)
ctx = host.TaaContext((W, H))
got = run_gpu_resolve(ctx, u, ins, hist, hist_depth=f0.depth.numpy(), want=want)
print("env", os.environ.get("TAA_STREAM_DEBUG"), "bad mask px:", int((ref["mask"] != got["mask"]).sum()))
for y in range(128, 144):
    print(y, "ref rejected cols:", int((ref["mask"][y] & 1).sum()), "got:", int((got["mask"][y] & 1).sum()),
          " max|d|:", float(np.abs(ref["history_out"][y].astype(np.float32) - got["history_out"][y].astype(np.float32)).max()))

for y in range(134, 144):
    print(y, " ".join(f"{got['mask'][y, x]:08x}" for x in (10, 11, 76, 77, 100, 200)))

import struct
for y in range(134, 144):
    a = int(got['mask'][y, 10]); b = int(got['mask'][y, 11])
    print(y, f"{a:08x}", struct.unpack('f', struct.pack('I', a))[0], f"{b:08x}")
