"""Input side: collate parity with the reference's collate_tts (golden fixture), pinned collator,
bucket sampler.  CPU only."""
import numpy as np
import pytest
import torch

from fastspeech2_b200.data import BucketBatchSampler, PinnedCollator, collate_tts

T_ = torch.from_numpy


def items_from(g):
    return [(g[f"x{i}"], g[f"mel{i}"], f"utt{i}", g[f"mel{i}"].shape[0], g[f"d{i}"], g[f"e{i}"], g[f"p{i}"]) for i in range(4)]


def check(out, g):
    for k, i in (("inputs", 0), ("ilens", 1), ("mels", 2), ("labels", 3), ("olens", 4), ("durations", 6), ("energys", 7), ("pitches", 8)):
        assert out[i].dtype == T_(g[k]).dtype and torch.equal(out[i], T_(g[k])), k
    assert out[5] == ["utt0", "utt1", "utt2", "utt3"]


def test_collate_matches_reference(golden):
    g = golden("collate")
    check(collate_tts(items_from(g)), g)


def test_pinned_collator_matches_and_reuses_buffers(golden):
    g = golden("collate")
    pc = PinnedCollator(max_batch=8, max_T=20, max_L=64, pin=False, slots=2)
    out = pc(items_from(g))
    check(out, g)
    ptr = out[2].data_ptr()
    mid = pc(items_from(g))                          # second slot: a different buffer set (the first may still be uploading)
    assert mid[2].data_ptr() != ptr and mid.slot != out.slot
    out2 = pc(items_from(g)[:2])                     # round robin back to the first slot, smaller batch: stale tail must be zero
    assert out2[2].data_ptr() == ptr and out2[2].shape == (2, 50, 80)
    ref2 = collate_tts(items_from(g)[:2])
    for i in (0, 1, 2, 3, 4, 6, 7, 8):
        assert torch.equal(out2[i], ref2[i])
    moved = pc.to_device(out2, "cpu")
    assert moved[5] == ["utt0", "utt1"] and torch.equal(moved[0], ref2[0])


def test_bucket_sampler_partitions_and_cuts_padding():
    rng = np.random.RandomState(0)
    lengths = (rng.gamma(4.0, 140.0, size=2000) + 100).astype(int)    # LJSpeech-like mel lengths
    s = BucketBatchSampler(lengths, batch_size=16, seed=3)
    batches = list(s)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(2000)) and len(batches) == len(s)
    assert list(BucketBatchSampler(lengths, 16, seed=3)) == batches               # deterministic per seed/epoch
    s.set_epoch(1)
    assert list(s) != batches
    rand_waste = 1.0 - float(np.mean(lengths)) / float(np.mean([lengths[np.array(b)].max() for b in
                        np.array_split(rng.permutation(2000), 125)]))
    assert s.padding_waste() < 0.1 < rand_waste


def test_vocoder_handoff_layout():
    """fastspeech2_b200.handoff vs the reference's own tensor ops (inference.py:170-180: audio.T, cat(dim=1), unsqueeze(0))."""
    from fastspeech2_b200.handoff import batch_to_vocoder, paragraph_mel, split_utterances
    g = torch.Generator().manual_seed(3)
    olens = torch.tensor([7, 3, 5])
    mels = torch.randn(3, 7, 80, generator=g)
    para_mel = []
    for b in range(3):
        audio = mels[b, :olens[b]]                 # what model.inference(text) returns: [L, 80]
        para_mel.append(audio.T)
    want = torch.cat(para_mel, dim=1).unsqueeze(0)  # inference.py:176,181
    assert want.shape == (1, 80, 15)
    assert torch.equal(batch_to_vocoder(mels, olens), want)
    assert torch.equal(paragraph_mel([mels[b, :olens[b]] for b in range(3)]), want)
    parts = split_utterances(mels, olens)
    assert [tuple(p.shape) for p in parts] == [(1, 80, 7), (1, 80, 3), (1, 80, 5)]
    assert torch.equal(torch.cat(parts, dim=2), want)


@pytest.mark.gpu
def test_pinned_collator_uploads_survive_run_ahead(golden):
    """a loader that runs ahead of the GPU must never refill a pinned slot an asynchronous upload is still
    reading.  Batches of different content are collated and uploaded back to back while the GPU is kept busy."""
    g = golden("collate")
    base = items_from(g)
    for slots in (1, 2):
        pc = PinnedCollator(max_batch=8, max_T=20, max_L=64, slots=slots)
        busy = torch.randn(4096, 4096, device="cuda")
        want, got = [], []
        for step in range(6):
            items = [(it[0], it[1] + step, it[2], it[3], it[4], it[5] * (step + 1), it[6]) for it in base]
            for _ in range(4):
                busy = busy @ busy * 1e-4                 # queue GPU work so the H2D copies lag behind the host
            col = pc(items)
            want.append(tuple(t.clone() if torch.is_tensor(t) else t for t in col))
            got.append(pc.to_device(col, "cuda"))
        torch.cuda.synchronize()
        for w, o in zip(want, got):
            for a, b in zip(w, o):
                if torch.is_tensor(a):
                    assert torch.equal(a, b.cpu())
