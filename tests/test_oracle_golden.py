"""Pin the CPU oracle (oracle/fs2_oracle.py) against vectors produced by the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import torch

from oracle import fs2_oracle as O
from conftest import GOLDEN

T = torch.from_numpy
FLOAT_TOL = 1e-6  # the same ATen ops in the same order; observed 0.0


def close(a, b, tol=FLOAT_TOL):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.max(np.abs(a - b)) <= tol, float(np.max(np.abs(a - b)))


def test_checkpoint_keys_match_reference(weights):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    assert [k for k, _, _ in ref] == list(weights.keys())
    for k, shape, dtype in ref:
        assert list(weights[k].shape) == shape, k
        assert str(weights[k].dtype) == dtype, k


def test_teacher_forced_ragged(golden, weights):
    g = golden("tf_ragged")
    torch.set_num_threads(1)
    with torch.no_grad():
        b, a, d, e, p = O.forward_path(weights, T(g["xs"]), T(g["ilens"]), T(g["olens"]), T(g["ds"]).clone(),
                                       T(g["es"]), T(g["ps"]), False)
    close(b, g["before"]); close(a, g["after"]); close(d, g["d_outs"]); close(e, g["e_outs"]); close(p, g["p_outs"])


def test_forward_loss(golden, weights):
    g, gl = golden("tf_ragged"), golden("tf_ragged_loss")
    with torch.no_grad():
        loss, report = O.forward_loss(weights, T(g["xs"]), T(g["ilens"]), T(g["ys"]), T(g["olens"]), T(g["ds"]).clone(),
                                      T(g["es"]), T(g["ps"]))
    keys = json.load(open(os.path.join(GOLDEN, "report_keys.json")))
    assert [list(r.keys())[0] for r in report] == keys
    close(loss, gl["loss"], 1e-5)
    close([list(r.values())[0] for r in report], gl["report"], 1e-5)


def test_unit_test_shapes(golden, weights):
    gl = golden("unit_shapes")
    x = torch.ones(2, 100, dtype=torch.int64); il = torch.tensor([100, 100])
    with torch.no_grad():
        loss, report = O.forward_loss(weights, x, il, torch.ones(2, 100, 80), il.clone(), torch.ones(2, 100),
                                      torch.ones(2, 100), torch.ones(2, 100))
    close(loss, gl["loss"], 1e-5)
    close([list(r.values())[0] for r in report], gl["report"], 1e-5)


def test_filelist_twin(golden, weights):
    """c2's ragged twin: 64 real filelist rows through the live reference, subset fixture."""
    from _synth import seeded_energy_pitch
    g = golden("filelist64")
    olens = T(g["olens"])
    L = int(olens.max())
    es, ps = seeded_energy_pitch(int(g["es_seed"]), olens, L)
    torch.set_num_threads(os.cpu_count())            # ~2.5 TFLOP of fp32; the summation order differs -> 1e-5, not 0
    with torch.no_grad():
        b, a, d, e, p = O.forward_path(weights, T(g["xs"]), T(g["ilens"]), olens, T(g["ds"]).clone(), es, ps, False)
    torch.set_num_threads(1)
    lo, hi = (int(v) for v in g["mel_rows"])
    close(d, g["d_outs"], 1e-5)
    close(e[T(g["ep_rows"])], g["e_sel"], 1e-5); close(p[T(g["ep_rows"])], g["p_sel"], 1e-5)
    close(a[lo, :olens[lo]], g["after_lo"], 1e-5); close(a[hi, :olens[hi]], g["after_hi"], 1e-5)
    close(b[lo, :olens[lo]], g["before_lo"], 1e-5)
    valid = (torch.arange(L)[None, :] < olens[:, None]).double()
    close((a.double() * valid[..., None]).sum((1, 2)) / (olens.double() * 80), g["after_mean"], 1e-6)
    close((b.double() * valid[..., None]).sum((1, 2)) / (olens.double() * 80), g["before_mean"], 1e-6)


def test_inference_ragged(golden, weights):
    g = golden("inf_ragged")
    torch.set_num_threads(1)
    with torch.no_grad():
        b, a, d, eh, ph = O.forward_path(weights, T(g["xs"]), T(g["ilens"]), is_inference=True)
    assert torch.equal(d, T(g["d_outs"]))
    assert torch.equal(eh.argmax(-1), T(g["e_ids"])) and torch.equal(ph.argmax(-1), T(g["p_ids"]))
    close(b, g["before"]); close(a, g["after"])


def test_inference_single(golden, weights):
    g = golden("inf_single")
    torch.set_num_threads(1)
    with torch.no_grad():
        mel = O.inference(weights, T(g["x"]))
    close(mel, g["mel"])


def test_length_regulator_bit_exact(golden):
    g = golden("length_regulator")
    hs, il = T(g["hs"]), T(g["ilens"])
    d = T(g["d_int"]).clone()
    assert torch.equal(O.length_regulator(hs, d, il), T(g["out_int"]))
    assert torch.equal(d, T(g["d_int_after"]))          # all-zero slice filled in place
    d = T(g["d_int"]).clone()
    assert torch.equal(O.length_regulator(hs, d, il, alpha=2.5), T(g["out_alpha"]))
    assert torch.equal(d, T(g["d_alpha_after"]))        # alpha != 1 works on a copy
    d = T(g["d_float"]).clone()
    assert torch.equal(O.length_regulator(hs, d, il), T(g["out_float"]))
    assert torch.equal(d, T(g["d_float_after"]))


def test_bucketize_edges(golden):
    g = golden("bucketize")
    assert torch.equal(O.bucket_ids(T(g["vals_e"]), T(g["e_bins"])), T(g["ids_e"]))
    assert torch.equal(O.bucket_ids(T(g["vals_p"]), T(g["p_bins"])), T(g["ids_p"]))


def test_variance_bins_match_reference(golden):
    from fastspeech2_b200.weights import ModelDims, variance_bins
    g = golden("bucketize")
    e, p = variance_bins(ModelDims())
    assert torch.equal(e, T(g["e_bins"])) and torch.equal(p, T(g["p_bins"]))
