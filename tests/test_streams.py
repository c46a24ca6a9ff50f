"""Batches of independent camera streams (taa_star_b200/streams.py, BASELINE configs[4])."""
import pytest
import torch

from taa_star_b200 import configs
from taa_star_b200.streams import streams_of_rank


def test_round_robin_deal_covers_every_stream_once():
    for world in (1, 2, 3, 4, 8):
        seen = []
        for r in range(world):
            mine = streams_of_rank(64, world, r)
            assert all(s % world == r for s in mine)
            seen += mine
        assert sorted(seen) == list(range(64))
    assert len(streams_of_rank(64, 8, 5)) == 8 and streams_of_rank(5, 8, 7) == []
    with pytest.raises(ValueError):
        streams_of_rank(64, 8, 8)


@pytest.mark.gpu
def test_batch_equals_one_context_per_stream_run_alone():
    """Frames of different camera streams interleaved on several CUDA streams give, per stream, bit for bit what the stream gives alone."""
    from taa_star_b200 import host
    from taa_star_b200.streams import StreamBatch
    from taa_star_b200.synth import SyntheticScene
    dev = torch.device("cuda")
    w, h, n, nframes = 256, 144, 5, 4
    scenes = [SyntheticScene(w, h, device=dev, with_aux=False, seed=100 + s, pan_px=(3.0 - s, 0.5 * s)) for s in range(n)]
    for p in (configs.config2_resolve(), configs.config3_full_chain()):
        frames = [[sc.frame(i) for i in range(nframes)] for sc in scenes]
        batch = StreamBatch(w, h, n, n_cuda_streams=3, device=dev)
        for i in range(nframes):
            for k in range(n):
                f = frames[k][i]
                u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(i == 0))
                batch.resolve(k, u, f.color, f.depth, f.velocity, history_depth=frames[k][max(i - 1, 0)].depth if p.mDepthCulling else None)
        torch.cuda.synchronize()
        for k in range(n):
            ctx = host.TaaContext((w, h))
            hist = [torch.zeros(h, w, 4, dtype=torch.float16, device=dev) for _ in range(2)]
            res = torch.zeros(h, w, 4, dtype=torch.float16, device=dev)
            for i in range(nframes):
                f = frames[k][i]
                u = configs.uniforms_for(p, f.jitter_ndc, reset_history=(i == 0))
                ctx.resolve(u, color=f.color, depth=f.depth, velocity=f.velocity, history_in=hist[i & 1], history_out=hist[1 - (i & 1)], result=res,
                            history_depth=frames[k][max(i - 1, 0)].depth if p.mDepthCulling else None)
            torch.cuda.synchronize()
            assert torch.equal(res.view(torch.int16), batch.result[k].view(torch.int16)), f"stream {k}: result differs"
            assert torch.equal(hist[nframes & 1].view(torch.int16), batch.hist[k][batch.parity[k]].view(torch.int16)), f"stream {k}: history differs"
            ctx.close()
        batch.close()
