"""The oracle restatement against the fixtures minted from the UNMODIFIED reference (oracle/make_golden.py)."""
import json
import os
import types

import numpy as np
import pytest
import torch

from lightningdot_b200 import synth
from oracle import evalloop, flatip, loss as oloss, towers


@pytest.mark.parametrize("name,kind,layers,seed,batch", [
    ("txt_l2", "txt", 2, 101, 6), ("img_l2", "img", 2, 102, 5), ("txt_l12", "txt", 12, 42, 4), ("img_l12", "img", 12, 42, 4)])
def test_tower_matches_reference(golden_dir, name, kind, layers, seed, batch):
    gold = np.load(os.path.join(golden_dir, f"tower_{name}.npz"))
    sd = synth.random_tower_state(kind, seed=seed, perturb=True, layers=layers)
    with torch.no_grad():
        if kind == "txt":
            b = synth.text_batch(batch, 32, seed=seed, ragged=True)
            seq, pooled = towers.text_tower(sd, b["input_ids"], b["attention_mask"], b["position_ids"])
        else:
            b = synth.image_batch(batch, 36, seed=seed, ragged=True)
            seq, pooled = towers.image_tower(sd, b["input_ids"], b["attention_mask"], b["position_ids"], b["img_feat"],
                                             b["img_pos_feat"], b["gather_index"])
    # fp32 tolerance: the reference itself moves by ~2e-6 with batch composition (hard part 2)
    np.testing.assert_allclose(pooled.numpy(), gold["pooled"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(seq[:, 0].numpy(), gold["cls_hidden"], atol=2e-5, rtol=0)


def test_inbatch_loss_matches_reference(golden_dir):
    gold = np.load(os.path.join(golden_dir, "loss_inbatch.npz"))
    g = torch.Generator().manual_seed(7)
    q = torch.randn(24, 768, generator=g) * 0.06
    ctx = torch.randn(24, 768, generator=g) * 0.06 + q * 0.45
    cap = torch.randn(24, 768, generator=g) * 0.06
    pos = list(range(24))
    l0, c0, s0 = oloss.nll(q, ctx, pos)
    l1, c1, s1 = oloss.nll(q, ctx, pos, cap, 0.1)
    assert abs(float(l0) - float(gold["loss0"])) < 1e-6 and int(c0) == int(gold["correct0"])
    assert abs(float(l1) - float(gold["loss1"])) < 1e-6 and int(c1) == int(gold["correct1"])
    np.testing.assert_allclose(s0.numpy(), gold["scores0"], atol=1e-6)
    np.testing.assert_allclose(s1.numpy(), gold["scores1"], atol=1e-6)


def test_flat_indexer_matches_reference_wrapper(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "indexer_small.json")))
    x = synth.gaussian_index(300, 768, seed=5)
    q, gt = synth.planted_queries(x, 17, sigma=1.0, seed=6)
    ids = [f"img_{i:07d}.npz" for i in range(300)]
    for scorer in (flatip.scores_f64, flatip.scores_f32):
        ix = flatip.FlatIndexer(768, buffer_size=128, scorer=scorer)
        ix.index_data(list(zip(ids, x)))
        res = ix.search_knn(q, 10)
        assert [list(r[0]) for r in res] == gold["ids"]
        np.testing.assert_allclose(np.stack([r[1] for r in res]), np.array(gold["scores"], np.float32), rtol=2e-6)
    assert [int(v) for v in gt] == gold["gt"]


def test_eval_loop_matches_reference(golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "evalloop_small.json")))
    n_img, cap_per_img = 200, 5
    x = synth.gaussian_index(n_img, 768, seed=11)
    n_cap = n_img * cap_per_img
    rng = np.random.default_rng(12)
    txt = (x[np.arange(n_cap) // cap_per_img] + 12.0 * rng.standard_normal((n_cap, 768), dtype=np.float32) / np.sqrt(768)).astype(np.float32)
    img = (x[np.arange(n_cap) // cap_per_img] + 1e-6 * rng.standard_normal((n_cap, 768), dtype=np.float32)).astype(np.float32)
    txt_ids = [str(j) for j in range(n_cap)]
    img_ids = [f"img_{j // cap_per_img:07d}.npz" for j in range(n_cap)]
    img2txt = {f"img_{i:07d}.npz": [str(i * cap_per_img + c) for c in range(cap_per_img)] for i in range(n_img)}
    rt, ri, rank_txt, rank_img = evalloop.recall_from_embeddings(txt, img, txt_ids, img_ids, img2txt, 100)
    assert {str(k): v for k, v in rt.items()} == gold["recall_txt"]
    assert {str(k): v for k, v in ri.items()} == gold["recall_img"]
    for k, v in gold["rank_txt_top10"].items():
        assert list(rank_txt[k][:10]) == v
    for k, v in gold["rank_img_top10"].items():
        assert list(rank_img[k][:10]) == v


def test_rank_rule_ties_and_short_index():
    # exact duplicates: equal scores are ordered by ascending row id; k > n pads with (-1, -FLT_MAX)
    x = np.tile(synth.gaussian_index(4, 64, seed=3), (3, 1))          # rows i, i+4, i+8 identical
    q = x[:2].copy()
    s, i = flatip.search(q, x, 16)
    assert i[0, :3].tolist() == [0, 4, 8] and i[1, :3].tolist() == [1, 5, 9]
    assert (i[:, 12:] == -1).all() and (s[:, 12:] == np.float32(-3.4028235e38)).all()
    assert (np.diff(s[:, :12].astype(np.float64), axis=1) <= 0).all()


def test_f64_oracle_vs_fp32_sgemm_agree_off_ties():
    """What faiss computes (fp32 sgemm) ranks like the fp64-accumulated bar except at near-ties (~1e-7 relative)."""
    x = synth.gaussian_index(20000, 768, seed=8)
    q, _ = synth.planted_queries(x, 64, sigma=3.0, seed=9)
    s64, i64 = flatip.search(q, x, 100, flatip.scores_f64)
    s32, i32 = flatip.search(q, x, 100, flatip.scores_f32)
    assert (i64 == i32).mean() > 0.999
    np.testing.assert_allclose(s64, s32, rtol=2e-5, atol=1e-7)
    diff_rows, diff_cols = np.nonzero(i64 != i32)
    for r, c in zip(diff_rows, diff_cols):  # every disagreement is a swap of two scores within fp32 noise
        assert abs(float(s64[r, c]) - float(s32[r, c])) <= 4e-7 * max(1.0, abs(float(s64[r, c])))


def test_oracle_train_step_matches_reference_gradients(golden_dir):
    """oracle/train.py (autograd over the functional towers) against the fixture minted from loss.backward() of the
    reference's own modules (oracle/make_golden.py: train_case)."""
    import torch
    from lightningdot_b200 import synth
    from oracle import train as otrain
    gold = np.load(os.path.join(golden_dir, "train_step_l2.npz"))
    layers, seed, batch = (int(v) for v in gold["meta"])
    sd_t = synth.random_tower_state("txt", seed=seed, perturb=True, layers=layers)
    sd_i = synth.random_tower_state("img", seed=seed + 1, perturb=True, layers=layers)
    tb = synth.text_batch(batch, 32, seed=seed, ragged=True)
    ib = synth.image_batch(batch, 36, seed=seed, ragged=True)
    loss, correct, gt, gi = otrain.train_step(sd_t, sd_i, tb, ib)
    assert abs(loss.item() - float(gold["loss"])) < 1e-5
    assert float(correct) == float(gold["correct"])
    for tag, grads in (("txt", gt), ("img", gi)):
        names = [str(n) for n in gold[f"{tag}_names"]]
        assert sorted(names) == sorted(grads.keys())
        for n, norm, samples in zip(names, gold[f"{tag}_norms"], gold[f"{tag}_samples"]):
            g = grads[n]
            assert abs(g.norm().item() - norm) <= 2e-4 * norm + 1e-6, n
            got = g.reshape(-1)[otrain.sample_index(g.numel())].numpy()
            assert np.abs(got - samples).max() <= 2e-4 * norm + 1e-7, n


def test_itm_batch_matches_reference_collate(golden_dir):
    """SURVEY 8 a1, the input contract: synth.itm_batch builds the nested batch the reference's own itm_fast_collate
    (dvl/data/itm.py:203-288) built from the same un-padded samples - every key, dtype, shape and byte, and the
    bookkeeping lists (tests/golden/itm_batch_schema.json, minted by oracle/make_golden.py)."""
    import json
    from lightningdot_b200 import synth
    gold = json.load(open(os.path.join(golden_dir, "itm_batch_schema.json")))
    for name, case in gold.items():
        got = synth.describe_batch(synth.itm_batch(**case["kwargs"]))
        assert got == case["batch"], name
        assert set(k.split(".")[0] for k in got) == {"txts", "imgs", "caps", "sample_size", "pos_ctx_indices",
                                                     "neg_ctx_indices", "txt_index", "img_fname"}
    # the per-sample tuples have the layout ItmFastDataset.__getitem__ produces (12 members, un-padded)
    s = synth.itm_samples(3, 16, 12, seed=1)
    assert len(s) == 3 and all(len(t) == 12 for t in s)
    ids, feat, pos, iid, m_t, m_i = s[0][:6]
    assert ids[0] == 101 and ids[-1] == 102 and m_t.shape == ids.shape and feat.shape[1] == 2048 and pos.shape[1] == 7
    assert iid.tolist() == [101] and m_i.shape[0] == feat.shape[0] + 1 and s[0][8] is None
