"""One row band of a larger frame on ONE GPU, no neighbours: what a rank's launch costs without any exchange (tuning aid for the unit geometry)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from taa_star_b200 import abi, configs, host
from taa_star_b200.synth import SyntheticScene
W, H = int(os.environ.get("W", 7680)), int(os.environ.get("H", 4320))
NB = int(os.environ.get("BANDS", 8)); R = int(os.environ.get("RANK_ID", 3)); halo = 20
y0, y1 = R * H // NB, (R + 1) * H // NB
hy0, hy1, iy0, iy1 = max(0, y0 - halo), min(H, y1 + halo), max(0, y0 - 2), min(H, y1 + 2)
dev = torch.device("cuda:0")
sc = SyntheticScene(W, H, device=dev, with_aux=False, rows=(iy0, iy1))
frames = [sc.frame(n) for n in range(4)]
p = configs.config2_resolve()
ctx = host.TaaContext((W, H), band=(y0, y1 - y0))
hist = [torch.zeros(hy1 - hy0, W, 4, dtype=torch.float16, device=dev) for _ in range(2)]
res = torch.zeros(y1 - y0, W, 4, dtype=torch.float16, device=dev)
if os.environ.get("PEER") == "1":
    # the band as its own neighbour on both sides (a torus): the PEER kernel variant with its dispatch order, second stores, waits and signals,
    # minus the NVLink hop — what the variant costs by itself
    import ctypes as C
    lib = abi.load_library()
    flags = torch.zeros(abi.TAA_BAND_FLAG_WORDS, dtype=torch.int32, device=dev)
    pb = abi.taa_band_peer()
    pb.history[0], pb.history[1] = hist[0].data_ptr(), hist[1].data_ptr()
    pb.row_pitch, pb.y0, pb.band_rows, pb.flags = W * 8, hy0, y1 - y0, flags.data_ptr()
    st = lib.taa_band_peers(ctx._h, C.byref(pb), C.byref(pb), hist[0].data_ptr(), hist[1].data_ptr(), flags.data_ptr(), halo)
    assert st == 0, lib.taa_last_error_string(ctx._h).decode()
stream = torch.cuda.Stream()
prep = []
for n in range(4):
    for par in range(2):
        f = frames[n]
        prep.append((ctx.images(color=(f.color, iy0), depth=(f.depth, iy0), velocity=(f.velocity, iy0), history_in=(hist[par], hy0), history_out=(hist[1 - par], hy0), result=(res, y0)),
                     configs.uniforms_for(p, f.jitter_ndc)))
def run(n):
    for i in range(n):
        im, u = prep[(i % 4) * 2 + (i % 2)]; ctx.resolve_prepared(im, u, stream.cuda_stream)
torch.cuda.synchronize()
run(8); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream); run(200); e1.record(stream); torch.cuda.synchronize()
print("band %dx%d of %dx%d: ms per frame %.5f  (peer=%s R=%s tail=%s rs=%s hints=%s) status %d" % (W, y1 - y0, W, H, e0.elapsed_time(e1) / 200, os.environ.get("PEER"), os.environ.get("TAA_STREAM_R"), os.environ.get("TAA_STREAM_TAIL"), os.environ.get("TAA_STREAM_RS"), os.environ.get("TAA_STREAM_HINTS"), ctx.poll_status(stream)))
