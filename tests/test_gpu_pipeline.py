"""GPU parity of the pieces around the two kernels families: in-batch NLL, the BiEncoder / eval-loop mirrors, the
multi-shard merge, and the launch accounting - all through the C ABI, checked against the oracle and the fixtures
minted from the reference."""
import json
import os
import types

import numpy as np
import pytest
import torch

from lightningdot_b200 import _lib, sharded, synth, trainer
from lightningdot_b200.bi_encoder import BertEncoder, BiEncoder, BiEncoderNllLoss, TowerConfig, \
    dot_product_scores
from lightningdot_b200.indexer import DenseFlatIndexer
from lightningdot_b200.utils import _calc_loss
from oracle import flatip, loss as oloss, towers as otowers

from test_host_logic import StubEncoder, check_evalloop_against_golden, evalloop_inputs

pytestmark = pytest.mark.gpu


def test_inbatch_nll_matches_reference_fixture(cuda_lib, golden_dir):
    """Tolerance: scores 1e-5 relative + 2e-6 absolute (fp16 hi/lo split products drop the lo*lo term, ~2^-21
    relative per operand pair, plus fp32 accumulation), loss 1e-5."""
    gold = np.load(os.path.join(golden_dir, "loss_inbatch.npz"))
    g = torch.Generator().manual_seed(7)
    q = torch.randn(24, 768, generator=g) * 0.06
    ctx = torch.randn(24, 768, generator=g) * 0.06 + q * 0.45
    cap = torch.randn(24, 768, generator=g) * 0.06
    pos = list(range(24))
    with torch.no_grad():
        l0, c0, s0 = _calc_loss(types.SimpleNamespace(caption_score_weight=0.0), BiEncoderNllLoss(), q.cuda(), ctx.cuda(),
                                None, pos, None)
        l1, c1, s1 = _calc_loss(types.SimpleNamespace(caption_score_weight=0.1), BiEncoderNllLoss(), q.cuda(), ctx.cuda(),
                                cap.cuda(), pos, None)
    assert abs(float(l0) - float(gold["loss0"])) < 1e-5 and int(c0) == int(gold["correct0"])
    assert abs(float(l1) - float(gold["loss1"])) < 1e-5 and int(c1) == int(gold["correct1"])
    np.testing.assert_allclose(s0.cpu().numpy(), gold["scores0"], atol=2e-6, rtol=1e-5)
    np.testing.assert_allclose(s1.cpu().numpy(), gold["scores1"], atol=2e-6, rtol=1e-5)


@pytest.mark.parametrize("bq,bc", [(1, 1), (7, 13), (96, 96), (300, 1000), (4096, 4096)])
def test_inbatch_nll_vs_oracle_shapes(cuda_lib, bq, bc):
    g = torch.Generator().manual_seed(bq * 31 + bc)
    q = torch.randn(bq, 768, generator=g) * 0.2
    ctx = torch.randn(bc, 768, generator=g) * 0.2
    pos = torch.randint(0, bc, (bq,), generator=g).tolist()
    with torch.no_grad():
        for red in ("mean", "sum"):
            l, c, s = BiEncoderNllLoss().calc(q.cuda(), ctx.cuda(), None, pos, reduction=red)
            ol, oc, os_ = oloss.nll(q.double(), ctx.double(), pos, reduction=red)
            assert abs(float(l) - float(ol)) <= 2e-5 * max(1.0, abs(float(ol)))
            assert int(c) == int(oc)
            assert torch.allclose(s.cpu().double(), os_, atol=5e-6, rtol=1e-5)
    # cosine flag of dot_product_scores (bi_encoder.py:63-67)
    cs = dot_product_scores(q.cuda(), ctx.cuda(), cosine=True).cpu()
    want = torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(ctx, dim=1).t()
    assert torch.allclose(cs, want, atol=1e-5)


def test_eval_loop_on_gpu_matches_reference_fixture(cuda_lib, golden_dir):
    """dvl/trainer.py:113-190 through the CUDA indexer + CUDA loss: recalls, ranked ids, loss and accuracy equal the
    fixture minted from the reference's own eval_model_on_dataloader."""
    gold = json.load(open(os.path.join(golden_dir, "evalloop_small.json")))
    txt, img, batches, img2txt = evalloop_inputs()
    args = types.SimpleNamespace(hnsw_index=False, vector_size=768, caption_score_weight=0.0)
    out = trainer.eval_model_on_dataloader(StubEncoder(txt, img, "cuda"), batches, args, img2txt, num_tops=100)
    # the in-batch accuracy is an argmax over scores that are tied to ~1e-6 relative in this fixture (every batch of
    # 16 holds several 1e-6-perturbed encodings of the same image), so it may move by a few samples with the
    # summation order of the score GEMM; recalls, ranks and the loss are compared exactly / to 1e-4
    check_evalloop_against_golden(out, gold, acc_tol=0.005)
    ix = trainer.get_indexer(StubEncoder(txt, img, "cuda"), batches, args, hnsw_index=False)
    assert ix.index_id_to_db_id == [f"img_{i:07d}.npz" for i in range(200)]
    assert np.array_equal(ix.index.vectors().cpu().numpy(), img[4::5])


def _small_biencoder(layers=2):
    args = types.SimpleNamespace(img_model_type='uniter-base', img_model_config=TowerConfig(num_hidden_layers=layers),
                                 img_checkpoint=None, txt_model_type='bert-base',
                                 txt_model_config=TowerConfig(num_hidden_layers=layers), txt_checkpoint=None)
    return BiEncoder(args, project_dim=768)


def test_biencoder_forward_on_collate_batch(cuda_lib):
    """BiEncoder.forward on the nested batch of itm_fast_collate (dvl/data/itm.py:203-288): txts / imgs / caps ->
    three pooled [B, 768] tensors, against the fp32 oracle towers.  Tolerance: bf16 towers, cosine >= 0.999."""
    model = _small_biencoder()
    sd_t = synth.random_tower_state("txt", seed=11, perturb=True, layers=2)
    sd_i = synth.random_tower_state("img", seed=12, perturb=True, layers=2)
    model.txt_model.load_state_dict(sd_t, strict=True)
    model.img_model.load_state_dict(sd_i, strict=True)
    assert set(model.state_dict().keys()) == {"txt_model." + k for k in sd_t} | {"img_model." + k for k in sd_i}
    model.cuda().eval()
    tb, ib = synth.text_batch(9, 32, seed=1, ragged=True), synth.image_batch(9, 36, seed=2, ragged=True)
    batch = {"txts": tb, "imgs": ib, "caps": synth.text_batch(9, 20, seed=3, ragged=True), "sample_size": 9}
    with torch.no_grad():
        t, i, c = model(batch)
        _, ot = otowers.text_tower(sd_t, tb["input_ids"], tb["attention_mask"], tb["position_ids"])
        _, oi = otowers.image_tower(sd_i, ib["input_ids"], ib["attention_mask"], ib["position_ids"], ib["img_feat"],
                                    ib["img_pos_feat"], ib["gather_index"])
        cb = batch["caps"]
        _, oc = otowers.text_tower(sd_t, cb["input_ids"], cb["attention_mask"], cb["position_ids"])
    for got, want in ((t, ot), (i, oi), (c, oc)):
        assert got.shape == want.shape and got.dtype == torch.float32
        assert torch.nn.functional.cosine_similarity(got.cpu(), want, dim=1).min() >= 0.999
    # caps absent -> None, sequence outputs on request
    with torch.no_grad():
        t2, i2, c2 = model({"txts": tb, "imgs": ib, "caps": {"input_ids": None}})
        seqs = model({"txts": tb, "imgs": ib}, output_all_encoded_layers=True)
    assert c2 is None and torch.equal(t2, t) and torch.equal(i2, i)
    assert seqs[0].shape == (9, 32, 768) and seqs[1].shape == (9, 37, 768) and seqs[2] is None


def test_sequence_outputs_carry_no_gradient(cuda_lib):
    """Gradients flow through the pooled outputs only (what train_itm.py uses); asking for differentiable sequence
    outputs fails loudly instead of returning a tensor that silently has no grad_fn."""
    model = _small_biencoder().cuda().train()
    tb = synth.text_batch(2, 16, seed=1)
    with pytest.raises(NotImplementedError):
        model({"txts": tb}, output_all_encoded_layers=True)


def test_towers_refuse_cpu():
    enc = BertEncoder(TowerConfig(num_hidden_layers=1), project_dim=768).eval()
    tb = synth.text_batch(2, 16, seed=1)
    with torch.no_grad(), pytest.raises(_lib.LdotError):
        enc(tb["input_ids"], tb["attention_mask"], tb["position_ids"])


@pytest.mark.parametrize("world,nq,k", [(2, 33, 10), (8, 200, 100), (3, 5, 1000)])
def test_topk_merge_kernel_vs_oracle(cuda_lib, world, nq, k):
    """Shard lists built by the oracle from a row-split index; merged on the GPU; equal to the single-index result."""
    n = 4000
    x = synth.gaussian_index(n, 64, seed=31)
    x[2500] = x[17]      # cross-shard duplicate
    q, _ = synth.planted_queries(x, nq, sigma=2.0, seed=32)
    b = sharded.shard_bounds(n, world)
    gs = np.empty((world, nq, k), np.float32)
    gi = np.empty((world, nq, k), np.int64)
    for r in range(world):
        s, i = flatip.search(q, x[b[r]:b[r + 1]], k)
        gs[r], gi[r] = s, np.where(i >= 0, i + b[r], -1)
    lib = cuda_lib
    ds, di = torch.from_numpy(gs).cuda(), torch.from_numpy(gi).cuda()
    out_s = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    out_i = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    _lib.check(lib.ldot_topk_merge(_lib.ptr(ds), _lib.ptr(di), world, nq, k, 0, 0, _lib.ptr(out_s), _lib.ptr(out_i),
                                   _lib.stream_ptr()))
    os_, oi = flatip.search(q, x, k)
    assert np.array_equal(out_i.cpu().numpy(), oi) and np.array_equal(out_s.cpu().numpy(), os_)


def test_sharded_indexer_single_rank_and_offsets(cuda_lib):
    """world = 1 path of ShardedFlatIndexer, and a hand-made 3-shard search (row_offset + merge) in one process."""
    n, k = 9000, 50
    x = synth.gaussian_index(n, 768, seed=41)
    q, _ = synth.planted_queries(x, 70, sigma=2.0, seed=42)
    ids = [f"d{i}" for i in range(n)]
    ix = sharded.ShardedFlatIndexer(768)
    ix.index_matrix(ids, x)
    res = ix.search_knn(q, k)
    os_, oi = flatip.search(q, x, k)
    assert [r[0] for r in res] == [[ids[j] for j in row] for row in oi]
    b = sharded.shard_bounds(n, 3)
    qd = torch.from_numpy(q).cuda()
    gs, gi = [], []
    for r in range(3):
        shard = sharded.FlatIPIndex(768, row_offset=b[r])
        shard.add(x[b[r]:b[r + 1]])
        s, i = shard.search_device(qd, k)
        gs.append(s)
        gi.append(i)
    ms, mi = ix._merge(torch.stack(gs), torch.stack(gi), k)
    assert np.array_equal(mi.cpu().numpy(), oi) and np.array_equal(ms.cpu().numpy(), os_)


def test_search_knn_accepts_device_tensors_and_remaps_ids(cuda_lib):
    x = synth.gaussian_index(50, 768, seed=51)
    ids = [("tuple", i) for i in range(50)]           # arbitrary hashable ids, as the reference allows
    ix = DenseFlatIndexer(768)
    ix.index_data(list(zip(ids, x)))
    a = ix.search_knn(x[:4], 60)                      # k > ntotal: label -1 maps to the LAST id (faiss_indexers.py:85)
    b = ix.search_knn(torch.from_numpy(x[:4]).cuda(), 60)
    assert [r[0] for r in a] == [r[0] for r in b]
    assert a[0][0][0] == ("tuple", 0) and a[0][0][50:] == [("tuple", 49)] * 10


def test_launch_accounting(cuda_lib):
    x = synth.gaussian_index(20000, 768, seed=61)
    q, _ = synth.planted_queries(x, 256, sigma=2.0, seed=62)
    ix = DenseFlatIndexer(768)
    ix.index_matrix(list(range(20000)), x)
    qd = torch.from_numpy(q).cuda()
    ix.index.search_device(qd, 100)
    torch.cuda.synchronize()
    _lib.prof_reset()
    _lib.prof_enable(True)
    for _ in range(3):
        ix.index.search_device(qd, 100)
    _lib.prof_enable(False)
    p = _lib.prof_read()
    c = p["coarse_score_topk"]
    assert c["launches"] == 3 and c["timed"] == 3 and c["ms"] > 0
    assert c["flops"] == 3 * 2.0 * 256 * 20000 * 768
    assert p["rescore"]["launches"] == 3 and p["select"]["launches"] >= 3 and p["linear_tcgen05"]["launches"] == 0


def test_hard_negative_mining_on_gpu(cuda_lib):
    """dvl/hn.py:47-68 through the CUDA eval loop with num_tops = 90 (num_hard_negatives = 40; the fixture's image index
    has 200 rows, and past the index size faiss pads with label -1, which the reference maps to the LAST db id): every
    sampled negative comes from the oracle's exact top-90 of its query and is never a positive."""
    import random
    from lightningdot_b200 import hn
    txt, img, batches, img2txt = evalloop_inputs()
    txt2img = {t: i for i, ts in img2txt.items() for t in ts}
    args = types.SimpleNamespace(hnsw_index=False, vector_size=768, caption_score_weight=0.0, num_hard_negatives=40)
    random.seed(3)
    neg_txt, neg_img = hn.sampled_hard_negatives(None, args, None, StubEncoder(txt, img, "cuda"), img2txt, txt2img,
                                                 train_dataloaders=[batches])
    assert len(neg_txt) == 200 and len(neg_img) == 1000
    k = hn.num_hard_sampled(40)
    assert k == 90
    _, oi = flatip.search(txt, img[4::5], k)          # text -> image (index rows = images in first-seen order)
    for j in (0, 17, 503, 999):
        got = neg_img[str(j)]
        want = {f"img_{r:07d}.npz" for r in oi[j]}
        assert len(got) == 40 and len(set(got)) == 40 and txt2img[str(j)] not in got and set(got) <= want
    _, oi2 = flatip.search(img[4::5], txt, k)         # image -> text
    for i in (0, 99, 199):
        name = f"img_{i:07d}.npz"
        got = neg_txt[name]
        assert len(got) == 40 and not set(got) & set(img2txt[name]) and set(got) <= {str(r) for r in oi2[i]}


def test_graphed_retriever_matches_eager_path(cuda_lib):
    """online.GraphedRetriever (token ids -> text tower -> exact top-k as one CUDA graph, rows a9 / f3) against the eager
    calls on the same padded inputs (bit-identical: same kernels, same data) and against the un-padded call of
    dvl/utils.py:204-211 (ids identical, scores to 1e-3 relative: padding changes only the softmax summation width)."""
    from lightningdot_b200.online import GraphedRetriever, retrieve_query
    sd = synth.random_tower_state("txt", seed=21, perturb=True, layers=2)
    model = BertEncoder(TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=2), project_dim=768)
    model.load_state_dict(sd, strict=True)
    model.cuda().eval()
    n, k, L = 6000, 10, 24
    tb = synth.text_batch(64, L, seed=5, ragged=True)
    with torch.no_grad():
        _, emb, _ = model(tb["input_ids"].cuda(), tb["attention_mask"].cuda(), tb["position_ids"].cuda(), need_sequence=False)
    x = synth.gaussian_index(n, 768, seed=9)
    x[100:164] = emb.cpu().numpy() * 0.9 + x[100:164] * 0.1      # rows close to the queries: a meaningful ranking
    ix = DenseFlatIndexer(768)
    ids = [f"img_{i:07d}.npz" for i in range(n)]
    ix.index_matrix(ids, x)
    gr = GraphedRetriever(model, ix, batch=4, seq_len=L, k=k)
    for lo in (0, 4, 9):
        nb = 3 if lo == 9 else 4
        q_ids, q_mask = tb["input_ids"][lo:lo + nb], tb["attention_mask"][lo:lo + nb]
        s, lab = gr.search(q_ids, q_mask)
        with torch.no_grad():
            _, e, _ = model(q_ids.cuda(), q_mask.cuda(), tb["position_ids"].cuda(), need_sequence=False)
        es, el = ix.index.search(e, k)
        assert np.array_equal(lab, el) and np.array_equal(s, es)
        os_, oi = flatip.search(e.cpu().numpy(), x, k)
        assert np.array_equal(lab, oi)
    assert gr.replays == 3 and gr.fallbacks == 0
    # one un-padded query, the retrieve_query way
    length = int(tb["attention_mask"][1].sum())
    tok = types.SimpleNamespace(encode=lambda text: tb["input_ids"][1, :length].tolist())
    res = retrieve_query(GraphedRetriever(model, ix, batch=1, seq_len=L, k=k), "a query", types.SimpleNamespace(tokenizer=tok))
    with torch.no_grad():
        _, e1, _ = model(tb["input_ids"][1:2, :length].cuda(), torch.ones(1, length, dtype=torch.long).cuda(),
                         torch.arange(length)[None].cuda(), need_sequence=False)
    want = ix.search_knn(e1, k)
    assert res[0][0] == want[0][0]
    np.testing.assert_allclose(res[0][1], want[0][1], rtol=1e-3)


def test_prefetch_loader_moves_nested_batches(cuda_lib):
    """loader.PrefetchLoader (uniter_model/data/loader.py mirror): nested itm_fast_collate batches arrive on the GPU in
    order, non-tensor members untouched, len() and attribute pass-through as the reference's."""
    from lightningdot_b200.loader import PrefetchLoader

    class Src(list):
        tag = "dataset-attr"
    batches = Src()
    for i in range(4):
        tb, ib = synth.text_batch(3, 8, seed=i), synth.image_batch(3, 5, seed=i)
        tb = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in tb.items()}
        batches.append({"txts": tb, "imgs": ib, "caps": {"input_ids": None}, "sample_size": 3, "txt_index": ["a", "b", "c"],
                        "pos_ctx_indices": [0, 1, 2]})
    pl = PrefetchLoader(batches)
    assert len(pl) == 4 and pl.tag == "dataset-attr"
    seen = 0
    for got, want in zip(pl, batches):
        assert got["txts"]["input_ids"].is_cuda and got["imgs"]["img_feat"].is_cuda
        assert torch.equal(got["txts"]["input_ids"].cpu(), want["txts"]["input_ids"])
        assert torch.equal(got["imgs"]["img_feat"].cpu(), want["imgs"]["img_feat"])
        assert got["caps"]["input_ids"] is None and got["txt_index"] == ["a", "b", "c"] and got["pos_ctx_indices"] == [0, 1, 2]
        seen += 1
    assert seen == 4


def test_full_pipeline_recall_vs_oracle_towers(cuda_lib):
    """BASELINE configs[0] in miniature, end to end: both CUDA towers (fp16, what `fp16: true` of the eval configs selects)
    -> eval_model_on_dataloader (encode, de-dup, index, two searches, Recall@1/5/10) against the fp32 CPU oracle towers ->
    oracle eval loop on the same inputs.  Random-init towers give nearly collinear embeddings, so ranks are decided by
    ~1e-3 relative score differences: the recalls must agree to +-0.01 absolute (they are near chance level anyway), the
    ranked top-10 lists must overlap by >= 97 % on average and >= 97 % of the top-1 hits must be the same image
    (measured on B200: recalls 0.006 / 0.052 / 0.104 vs 0.006 / 0.052 / 0.102 text->image, identical image->text;
    top-10 overlap 0.994 / 0.991; top-1 equal 0.992)."""
    from oracle import evalloop
    layers, n_img, cap_per_img, bs, L, R = 4, 100, 5, 50, 24, 20
    n_cap = n_img * cap_per_img
    sd_t = synth.random_tower_state("txt", seed=31, perturb=True, layers=layers)
    sd_i = synth.random_tower_state("img", seed=32, perturb=True, layers=layers)
    model = _small_biencoder(layers)
    model.txt_model.load_state_dict(sd_t, strict=True)
    model.img_model.load_state_dict(sd_i, strict=True)
    model.cuda().eval()
    model.txt_model.compute_dtype = model.img_model.compute_dtype = torch.float16
    tb = synth.text_batch(n_cap, L, seed=7, ragged=True)
    ib = synth.image_batch(n_img, R, seed=8, ragged=True)
    txt_ids = [str(j) for j in range(n_cap)]
    img_ids = [f"img_{j // cap_per_img:07d}.npz" for j in range(n_cap)]
    img2txt = {f"img_{i:07d}.npz": [str(i * cap_per_img + c) for c in range(cap_per_img)] for i in range(n_img)}
    batches = []
    for b0 in range(0, n_cap, bs):
        rows = torch.arange(b0, b0 + bs)
        irows = rows // cap_per_img
        batches.append({"txts": {k: (v[rows] if torch.is_tensor(v) and v.shape[0] == n_cap else v) for k, v in tb.items()},
                        "imgs": {k: (v[irows] if torch.is_tensor(v) and v.shape[0] == n_img else v) for k, v in ib.items()},
                        "caps": {"input_ids": None}, "sample_size": bs, "txt_index": txt_ids[b0:b0 + bs],
                        "img_fname": img_ids[b0:b0 + bs]})
    args = types.SimpleNamespace(hnsw_index=False, vector_size=768, caption_score_weight=0.0)
    _, _, _, (r_txt, r_img), (rank_txt, rank_img) = trainer.eval_model_on_dataloader(model, batches, args, img2txt, 50)
    with torch.no_grad():
        _, ot = otowers.text_tower(sd_t, tb["input_ids"], tb["attention_mask"], tb["position_ids"])
        _, oi = otowers.image_tower(sd_i, ib["input_ids"], ib["attention_mask"], ib["position_ids"], ib["img_feat"],
                                    ib["img_pos_feat"], ib["gather_index"])
    o_img = oi.numpy()[np.arange(n_cap) // cap_per_img]
    or_txt, or_img, orank_txt, orank_img = evalloop.recall_from_embeddings(ot.numpy(), o_img, txt_ids, img_ids, img2txt, 50)
    ov_txt = np.mean([len(set(rank_txt[q][:10]) & set(orank_txt[q][:10])) / 10 for q in txt_ids])
    ov_img = np.mean([len(set(rank_img[q][:10]) & set(orank_img[q][:10])) / 10 for q in img2txt])
    top1 = np.mean([rank_txt[q][0] == orank_txt[q][0] for q in txt_ids])
    msg = f"recall txt {r_txt} vs {or_txt}; img {r_img} vs {or_img}; top-10 overlap {ov_txt:.3f} / {ov_img:.3f}; top-1 equal {top1:.3f}"
    print(msg)
    for k in (1, 5, 10):
        assert abs(r_txt[k] - or_txt[k]) <= 0.01 and abs(r_img[k] - or_img[k]) <= 0.01, msg
    assert ov_txt >= 0.97 and ov_img >= 0.97 and top1 >= 0.97, msg
