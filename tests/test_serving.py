"""TorchScript / serving variant (reference: utils/fastspeech2_script.py + export_torchscript.py):
the exported module is scriptable, survives save / load, and its graph is one call of the custom operator that runs
libfs2b200.so.  The CPU part checks the host logic; execution needs a GPU."""
import pytest
import torch

from fastspeech2_b200 import FeedForwardTransformer
from fastspeech2_b200.hparams import load_hp
from fastspeech2_b200.serving import export_torchscript, pack_state, scripted, unpack_state


def _model(weights, precision=None):
    m = FeedForwardTransformer(68, 80, load_hp(), precision=precision)
    m.load_state_dict(weights, strict=True)
    return m.eval()


def test_checkpoint_blob_round_trip(weights):
    blob, keys, ranks, dims = pack_state(weights)
    back = unpack_state(blob, keys, ranks, dims)
    assert list(back) == list(weights)
    assert all(back[k].dtype == weights[k].dtype and torch.equal(back[k], weights[k]) for k in weights)


def test_script_save_load_and_loud_cpu(tmp_path, weights):
    path = export_torchscript(_model(weights), str(tmp_path / "fs2.pt"))
    served = torch.jit.load(path)
    assert "fs2_b200::inference" in str(served.graph)                      # the whole forward is the custom operator
    assert hasattr(served, "batch")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        served(torch.ones(5, dtype=torch.int64))


@pytest.mark.gpu
def test_served_module_matches_model_on_gpu(tmp_path, weights):
    m = _model(weights).cuda()
    served = torch.jit.load(export_torchscript(m, str(tmp_path / "fs2.pt"))).cuda()
    g = torch.Generator().manual_seed(9)
    x = torch.randint(1, 68, (37,), generator=g).cuda()
    with torch.no_grad():
        want = m.inference(x)
    got = served(x)
    assert got.shape == want.shape and torch.equal(got, want)              # same kernels, same weights -> bit-identical
    xs = torch.zeros(3, 37, dtype=torch.int64); il = torch.tensor([37, 20, 5])
    for b, n in enumerate(il.tolist()):
        xs[b, :n] = torch.randint(1, 68, (n,), generator=g)
    mels, olens = served.batch(xs.cuda(), il.cuda())
    with torch.no_grad():
        _, after, d, _, _ = m._forward(xs.cuda(), il.cuda(), is_inference=True, _one_hot=False)
    assert torch.equal(mels, after) and torch.equal(olens, d.sum(1))
    traced = torch.jit.trace(scripted(m), x)                              # export_torchscript.py:51-57 (--trace)
    assert torch.equal(traced(x), want)
