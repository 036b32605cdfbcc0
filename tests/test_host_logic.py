"""Host-side logic that needs no GPU: the eval loop mirror (dvl/trainer.py:113-190) against the fixture minted from
the reference, shard arithmetic, and the row-sharded search protocol over a 2-rank gloo group (device kernels
replaced by the oracle so that the exchange / offset / merge plumbing is what is tested)."""
import json
import os
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightningdot_b200 import sharded, synth, trainer
from oracle import flatip, loss as oloss


class OracleIndexer:
    """CPU stand-in with the DenseFlatIndexer surface the eval loop touches."""

    def __init__(self, vector_sz, **kw):
        self.inner = flatip.FlatIndexer(vector_sz)
        self.index_id_to_db_id = self.inner.index_id_to_db_id

    def index_matrix(self, ids, vectors):
        self.inner.index_data(list(zip(ids, vectors.cpu().numpy())))

    def search_knn(self, q, k):
        return self.inner.search_knn(q.cpu().numpy() if isinstance(q, torch.Tensor) else q, k)


class OracleLoss:
    def calc(self, q, ctx, cap, pos, hard=None, caption_score_weight=0.1, experiment=None, reduction='mean'):
        return oloss.nll(q, ctx, pos, cap, caption_score_weight, reduction)


def evalloop_inputs():
    n_img, cap_per_img, bs = 200, 5, 16
    x = synth.gaussian_index(n_img, 768, seed=11)
    n_cap = n_img * cap_per_img
    rng = np.random.default_rng(12)
    txt = (x[np.arange(n_cap) // cap_per_img] + 12.0 * rng.standard_normal((n_cap, 768), dtype=np.float32) / np.sqrt(768)).astype(np.float32)
    img = (x[np.arange(n_cap) // cap_per_img] + 1e-6 * rng.standard_normal((n_cap, 768), dtype=np.float32)).astype(np.float32)
    txt_ids = [str(j) for j in range(n_cap)]
    img_ids = [f"img_{j // cap_per_img:07d}.npz" for j in range(n_cap)]
    img2txt = {f"img_{i:07d}.npz": [str(i * cap_per_img + c) for c in range(cap_per_img)] for i in range(n_img)}
    batches = []
    for b in range(0, n_cap, bs):
        sl = slice(b, min(n_cap, b + bs))
        batches.append({"txts": {"input_ids": torch.zeros(sl.stop - sl.start, 4, dtype=torch.long)},
                        "txt_index": txt_ids[sl], "img_fname": img_ids[sl], "_slice": sl})
    return txt, img, batches, img2txt


class StubEncoder:
    def __init__(self, txt, img, device="cpu"):
        self.txt, self.img, self.device = txt, img, device

    def eval(self):
        return self

    def __call__(self, batch):
        sl = batch["_slice"]
        return (torch.from_numpy(self.txt[sl]).to(self.device), torch.from_numpy(self.img[sl]).to(self.device), None)


def check_evalloop_against_golden(out, gold, acc_tol=1e-9):
    loss, acc, (ix_img, ix_txt), (recall_txt, recall_img), (rank_txt, rank_img) = out
    assert {str(k): v for k, v in recall_txt.items()} == gold["recall_txt"]
    assert {str(k): v for k, v in recall_img.items()} == gold["recall_img"]
    assert abs(loss - gold["loss"]) < 1e-4 and abs(acc - gold["acc"]) <= acc_tol
    for k, v in gold["rank_txt_top10"].items():
        assert list(rank_txt[k][:10]) == v
    for k, v in gold["rank_img_top10"].items():
        assert list(rank_img[k][:10]) == v
    assert len(ix_img.index_id_to_db_id) == 200 and len(ix_txt.index_id_to_db_id) == 1000


def test_eval_loop_mirror_matches_reference_fixture(golden_dir, monkeypatch):
    gold = json.load(open(os.path.join(golden_dir, "evalloop_small.json")))
    txt, img, batches, img2txt = evalloop_inputs()
    monkeypatch.setattr(trainer, "DenseFlatIndexer", OracleIndexer)
    monkeypatch.setattr(trainer, "BiEncoderNllLoss", OracleLoss)
    args = types.SimpleNamespace(hnsw_index=False, vector_size=768, caption_score_weight=0.0)
    out = trainer.eval_model_on_dataloader(StubEncoder(txt, img), batches, args, img2txt, num_tops=100)
    check_evalloop_against_golden(out, gold)
    # no_eval: indexes only
    out2 = trainer.eval_model_on_dataloader(StubEncoder(txt, img), batches, args, img2txt, no_eval=True)
    assert out2[3] == (None, None) and out2[4] == (None, None)


def test_get_indexer_keeps_last_encoding_in_first_seen_order(monkeypatch):
    txt, img, batches, _ = evalloop_inputs()
    monkeypatch.setattr(trainer, "DenseFlatIndexer", OracleIndexer)
    args = types.SimpleNamespace(vector_size=768)
    ix = trainer.get_indexer(StubEncoder(txt, img), batches, args, hnsw_index=False, img_retrieval=True)
    assert ix.index_id_to_db_id == [f"img_{i:07d}.npz" for i in range(200)]
    # image i was encoded by captions 5i .. 5i+4: the LAST one (row 5i + 4) is what the index holds (trainer.py:151)
    assert np.array_equal(ix.inner.xb, img[4::5])


def test_checkpoint_state_roundtrip(tmp_path):
    model = torch.nn.Linear(4, 3)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    args = types.SimpleNamespace(output_dir=str(tmp_path))
    cp = trainer._save_checkpoint(args, model, opt, sched, epoch=3, offset=0, cp_name="best")
    assert cp.endswith("biencoder.best.pt")
    assert trainer._save_checkpoint(args, model, opt, sched, epoch=3, offset=7).endswith("biencoder.3.7.pt")
    st = trainer.load_states_from_checkpoint(cp)
    assert set(st._fields) == {'model_dict', 'optimizer_dict', 'scheduler_dict', 'offset', 'epoch', 'encoder_params'}
    model2 = torch.nn.Linear(4, 3)
    trainer.load_saved_state(model2, saved_state=st)
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), model2.state_dict().values()))


def test_shard_bounds():
    assert sharded.shard_bounds(1_000_000, 8) == [i * 125000 for i in range(9)]
    assert sharded.shard_bounds(10, 4) == [0, 3, 6, 8, 10]
    assert sharded.shard_bounds(2, 4) == [0, 1, 2, 2, 2]
    b = sharded.shard_bounds(123287, 8)
    assert b[0] == 0 and b[-1] == 123287 and max(np.diff(b)) - min(np.diff(b)) <= 1


class _OracleLocalIndex:
    """FlatIPIndex stand-in on CPU tensors (row_offset semantics included)."""

    def __init__(self, d, row_offset=0, **kw):
        self.d, self.row_offset, self.x = d, row_offset, np.zeros((0, d), np.float32)
        self.last_flagged = 0

    @property
    def ntotal(self):
        return len(self.x)

    def add(self, v):
        v = v.numpy() if isinstance(v, torch.Tensor) else v
        self.x = np.concatenate([self.x, np.asarray(v, np.float32)], 0)

    def search_device(self, q, k):
        s, i = flatip.search(q.numpy(), self.x, k)
        i = np.where(i >= 0, i + self.row_offset, -1)
        return torch.from_numpy(s), torch.from_numpy(i)

    def _device(self):
        return torch.device("cpu")

    def to_host(self, scores, idx):
        return scores.numpy(), idx.numpy()


class _CpuSharded(sharded.ShardedFlatIndexer):
    def _make_local_index(self, row_offset):
        return _OracleLocalIndex(self.vector_sz, row_offset=row_offset)

    def _local_search(self, queries, k, defer=False):
        s, i = self.index.search_device(queries, k) if self.index.ntotal else \
            (torch.full((queries.shape[0], k), -3.4028235e38), torch.full((queries.shape[0], k), -1, dtype=torch.int64))
        return (s, i, torch.zeros(1, dtype=torch.int32)) if defer else (s, i)

    def _merge_packed(self, recv, W, m, k, out):
        # the packed exchange layout of sharded.py: per shard (scores fp32 [m, k] | ids int64 [m, k]) as bytes
        gs = recv[:, :4 * m * k].contiguous().view(torch.float32).view(W, m, k)
        gi = recv[:, 4 * m * k:].contiguous().view(torch.int64).view(W, m, k)
        s, i = self._merge(gs, gi, k)
        out[:4 * m * k].view(torch.float32).copy_(s.reshape(-1))
        out[4 * m * k:12 * m * k].view(torch.int64).copy_(i.reshape(-1))

    def _merge(self, gs, gi, k):
        # (score desc, id asc) over the W * k gathered candidates of every query - what ldot_topk_merge does
        world, nq, _ = gs.shape
        s = gs.permute(1, 0, 2).reshape(nq, world * k).numpy()
        i = gi.permute(1, 0, 2).reshape(nq, world * k).numpy()
        out_s = np.full((nq, k), np.float32(-3.4028235e38), np.float32)
        out_i = np.full((nq, k), -1, np.int64)
        for r in range(nq):
            valid = np.nonzero(i[r] >= 0)[0]
            order = valid[np.lexsort((i[r, valid], -s[r, valid].astype(np.float64)))][:k]
            out_s[r, :len(order)] = s[r, order]
            out_i[r, :len(order)] = i[r, order]
        return torch.from_numpy(out_s), torch.from_numpy(out_i)


def _sharded_worker(rank, world, port, n, nq, k, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = synth.gaussian_index(n, 64, seed=21)
        if n > 16:
            x[n // 2 + 3] = x[5]        # a cross-shard exact duplicate: tie broken by the global row id
        q, _ = synth.planted_queries(x, nq, sigma=2.0, seed=22)
        ids = [f"doc{i}" for i in range(n)]
        ix = _CpuSharded(64)
        ix.index_matrix(ids, x)
        assert ix.index.ntotal == ix.bounds[rank + 1] - ix.bounds[rank] and ix.index.row_offset == ix.bounds[rank]
        # queries arrive sharded (each rank "encoded" its slice) and are gathered in rank order
        qb = sharded.shard_bounds(nq, world)
        q_all = ix.gather_queries(torch.from_numpy(q[qb[rank]:qb[rank + 1]]))
        assert np.array_equal(q_all.numpy(), q)
        # caller-supplied counts (no count exchange / host sync): ragged blocks, and equal blocks in one collective
        counts = [qb[r + 1] - qb[r] for r in range(world)]
        assert np.array_equal(ix.gather_queries(torch.from_numpy(q[qb[rank]:qb[rank + 1]]), counts).numpy(), q)
        even = (nq // world) * world
        eb = even // world
        if eb:
            got = ix.gather_queries(torch.from_numpy(q[rank * eb:(rank + 1) * eb]), [eb] * world)
            assert np.array_equal(got.numpy(), q[:even])
        res = ix.search_knn(q_all, k)
        os_, oi = flatip.search(q, x, k)
        want = [[ids[j] if j >= 0 else ids[-1] for j in row] for row in oi]
        assert [r[0] for r in res] == want
        assert np.array_equal(np.stack([r[1] for r in res]), os_)
        # already-partitioned build path
        ix2 = _CpuSharded(64)
        ix2.index_shard(ids, x[ix.bounds[rank]:ix.bounds[rank + 1]], ix.bounds)
        s2, i2 = ix2.search_device(q_all, k)
        assert np.array_equal(i2.numpy(), oi) and np.array_equal(s2.numpy(), os_)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,nq,k", [(1001, 37, 10), (3, 5, 4)])
def test_sharded_search_protocol_two_ranks_gloo(tmp_path, n, nq, k):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker, args=(2, port, n, nq, k, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _eval_loop_worker(rank, world, port, gold_path, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lightningdot_b200.sharded as sh
        gold = json.load(open(gold_path))
        txt, img, _, img2txt = evalloop_inputs()
        n_cap = len(txt)
        txt_ids = [str(j) for j in range(n_cap)]
        img_ids = [f"img_{j // 5:07d}.npz" for j in range(n_cap)]
        mine = np.arange(rank, n_cap, world)          # TxtTokLmdb's ids[rank::world]
        batches = [{"txts": {"input_ids": torch.zeros(len(rows), 4, dtype=torch.long)},
                    "txt_index": [txt_ids[j] for j in rows], "img_fname": [img_ids[j] for j in rows], "_rows": rows}
                   for rows in np.array_split(mine, 32)]

        class StridedEncoder:
            def eval(self):
                return self

            def __call__(self, batch):
                return torch.from_numpy(txt[batch["_rows"]]), torch.from_numpy(img[batch["_rows"]]), None

        trainer.BiEncoderNllLoss = OracleLoss
        sh.ShardedFlatIndexer = _CpuSharded           # (trainer._new_indexer imports it from the module at call time)
        args = types.SimpleNamespace(hnsw_index=False, vector_size=768, caption_score_weight=0.0)
        out = trainer.eval_model_on_dataloader(StridedEncoder(), batches, args, img2txt, num_tops=100)
        loss, acc, (ix_img, ix_txt), (recall_txt, recall_img), (rank_txt, rank_img) = out
        assert isinstance(ix_img, _CpuSharded) and ix_img.index.ntotal == 100 and ix_txt.index.ntotal == 500
        assert {str(k): v for k, v in recall_txt.items()} == gold["recall_txt"]
        assert {str(k): v for k, v in recall_img.items()} == gold["recall_img"]
        assert all(list(rank_txt[k][:10]) == v for k, v in gold["rank_txt_top10"].items())
        assert all(list(rank_img[k][:10]) == v for k, v in gold["rank_img_top10"].items())
        assert ix_img.index_id_to_db_id == [f"img_{i:07d}.npz" for i in range(200)] and ix_txt.index_id_to_db_id == txt_ids
        assert np.isfinite(loss) and 0.0 <= acc <= 1.0
        ix = trainer.get_indexer(StridedEncoder(), batches, args, hnsw_index=False, img_retrieval=True)
        assert ix.index_id_to_db_id == [f"img_{i:07d}.npz" for i in range(200)] and ix.n_global == 200
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_eval_loop_two_ranks_gloo_equals_single_process_fixture(golden_dir, tmp_path):
    """eval_model_on_dataloader under a process group: every rank evaluates the captions ids[rank::world], embeddings are
    pooled into the single-process order, the indexes are row-sharded - recalls and rankings equal the fixture minted
    from the reference's single-process loop, on every rank."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_eval_loop_worker, args=(2, port, os.path.join(golden_dir, "evalloop_small.json"), str(tmp_path)), nprocs=2,
             join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _loss_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightningdot_b200 import utils
        g = torch.Generator().manual_seed(5)
        q = torch.randn(16, 32, generator=g)
        ctx = torch.randn(16, 32, generator=g) + q
        lo, hi = rank * 8, rank * 8 + 8
        args = types.SimpleNamespace(distributed_world_size=world, caption_score_weight=0.0)
        loss, correct, scores = utils._calc_loss(args, OracleLoss(), q[lo:hi], ctx[lo:hi], None, list(range(8)), None)
        # the global-batch loss restricted to this rank's query rows
        want_l, want_c, want_s = oloss.nll(q[lo:hi], ctx, list(range(lo, hi)))
        assert torch.allclose(loss, want_l) and int(correct) == int(want_c) and torch.allclose(scores, want_s)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_inbatch_loss_global_negatives_two_ranks_gloo(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_loss_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def _train_sync_worker(rank, world, port, tmp):
    """Global-batch training protocol: differentiable embedding gather + gradient average over the ranks
    must reproduce the single-process gradient of the symmetric loss on the whole batch.  Towers are stand-in
    torch Linears (the protocol is host logic; the CUDA towers are tested on the GPU)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lightningdot_b200 import utils
        g = torch.Generator().manual_seed(9)
        B, b = 12, 12 // world
        xt, xi = torch.randn(B, 20, generator=g), torch.randn(B, 24, generator=g)

        def towers():
            torch.manual_seed(3)
            return torch.nn.Linear(20, 16), torch.nn.Linear(24, 16)

        # single process, global batch
        ft, fi = towers()
        loss_ref, _ = oloss.symmetric_nll(ft(xt), fi(xi))
        loss_ref.backward()
        # this rank's slice
        mt, mi = towers()
        lo, hi = rank * b, rank * b + b
        t, i = mt(xt[lo:hi]), mi(xi[lo:hi])
        args = types.SimpleNamespace(distributed_world_size=world, caption_score_weight=0.0)
        pos = list(range(b))
        l_txt, _, _ = utils._calc_loss(args, OracleLoss(), i, t, None, pos, None)
        l_img, _, _ = utils._calc_loss(args, OracleLoss(), t, i, None, pos, None)
        loss = 0.5 * l_txt + 0.5 * l_img
        loss.backward()
        params = list(mt.parameters()) + list(mi.parameters())
        utils.sync_gradients(params)
        for p, r in zip(params, list(ft.parameters()) + list(fi.parameters())):
            assert torch.allclose(p.grad, r.grad, rtol=1e-4, atol=1e-6), (p.grad - r.grad).abs().max()
        mean_loss = utils._mean_or_sum_(loss.detach().clone(), None, mean=True)
        assert torch.allclose(mean_loss, loss_ref.detach(), rtol=1e-5)
        # no-grad gather path returns the same rows
        with torch.no_grad():
            assert torch.equal(utils.gather_embeddings(t.detach()), ft(xt).detach()) or \
                torch.allclose(utils.gather_embeddings(t.detach()), ft(xt).detach())
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_global_batch_training_protocol_two_ranks_gloo(tmp_path):
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_train_sync_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(tmp_path / "ok0") and os.path.exists(tmp_path / "ok1")


def test_hard_negative_sampling_logic(tmp_path, monkeypatch):
    """dvl/hn.py: candidate count formula, positives removed (image side in place / order kept, text side through a set),
    num_hard_negatives drawn without replacement, per-dataset results chained; mapping builder on img2txts.json."""
    import json as _json
    import random as _random
    from lightningdot_b200 import hn
    assert [hn.num_hard_sampled(k) for k in (0, 10, 20, 100, 600)] == [50, 50, 50, 210, 1000]
    img2txt = {"a.npz": ["0", "1"], "b.npz": ["2", "3"], "c.npz": ["4"]}
    txt2img = {t: i for i, ts in img2txt.items() for t in ts}
    rank_txt = {"0": ["b.npz", "a.npz", "c.npz"], "2": ["a.npz", "c.npz", "b.npz"], "4": ["a.npz", "b.npz"]}
    rank_img = {"a.npz": ["1", "2", "0", "4"], "b.npz": ["4", "0", "3", "1"], "c.npz": ["0", "2", "4"]}
    _random.seed(0)
    t_out, i_out = hn.filter_and_sample({k: list(v) for k, v in rank_txt.items()}, rank_img, img2txt, txt2img, 2)
    assert set(t_out) == set(rank_img) and set(i_out) == set(rank_txt)
    for img, negs in t_out.items():
        assert len(negs) == 2 and len(set(negs)) == 2 and not set(negs) & set(img2txt[img]) and set(negs) <= set(rank_img[img])
    for txt, negs in i_out.items():
        assert len(negs) == 2 and txt2img[txt] not in negs and set(negs) <= set(rank_txt[txt])
    # plumbing: one eval per dataloader with k = num_hard_sampled, later datasets do not override earlier keys (ChainMap)
    calls = []

    def fake_eval(model, loader, args, i2t, k):
        calls.append((loader, k))
        return 0.0, 0.0, None, None, ({k_: list(v) for k_, v in rank_txt.items()}, dict(rank_img))
    monkeypatch.setattr(hn, "eval_model_on_dataloader", fake_eval)
    args = types.SimpleNamespace(num_hard_negatives=2)
    t_all, i_all = hn.sampled_hard_negatives(None, args, None, object(), img2txt, txt2img, train_dataloaders=[[1], [2]])
    assert calls == [([1], 50), ([2], 50)] and set(t_all) == set(rank_img) and set(i_all) == set(rank_txt)
    # default path: the per-dataset loaders are built from the database folders as dvl/hn.py:48-52 does
    from lightningdot_b200 import data as mdata
    txt_dir, img_dir = synth.make_itm_db(str(tmp_path / "db"), 6, 2, seq_len=16, num_bb=12, seed=4)
    built = []
    monkeypatch.setattr(hn, "build_dataloader", lambda dset, collate, is_train, a, bs=None: built.append((dset, is_train, bs)) or [len(dset)])
    args = types.SimpleNamespace(num_hard_negatives=2, train_txt_dbs=[txt_dir], train_img_dbs=[img_dir], max_txt_len=60,
                                 img_meta=None, tokenizer=None, valid_batch_size=5)
    calls.clear()
    hn.sampled_hard_negatives(mdata.ImageLmdbGroup(0.2, 100, 10, 36, False), args, None, object(), img2txt, txt2img)
    assert calls == [([12], 50)] and isinstance(built[0][0], mdata.ItmFastDataset) and built[0][1:] == (True, 5)
    assert built[0][0].train_imgs is not None and built[0][0].neg_imgs == [None] * 12      # new_epoch() without negatives
    # mappings
    for name, part in (("d1", {"a.npz": ["0", "1"]}), ("d2", {"b.npz": ["2", "3"], "c.npz": ["4"]})):
        os.makedirs(tmp_path / name)
        _json.dump(part, open(tmp_path / name / "img2txts.json", "w"))
    dbs = [str(tmp_path / "d1"), str(tmp_path / "d2")]
    i2t, t2i, i2s, t2s, s2i, s2t = hn.get_img_txt_mappings(dbs)
    assert i2t == img2txt and t2i == txt2img and i2s["c.npz"] == dbs[1] and t2s["1"] == dbs[0]
    assert sorted(s2i[dbs[1]]) == ["b.npz", "c.npz"] and sorted(s2t[dbs[1]]) == ["2", "3", "4"]
    # an image listed by two folders belongs to the FIRST one (the reference resolves through a ChainMap); expected values
    # below were produced by the reference's own get_img_txt_mappings on the same files
    for name, part in (("e0", {"a": ["0", "1"], "b": ["2"]}), ("e1", {"b": ["9"], "c": ["3", "4"]}), ("e2", {"d": ["5"]})):
        os.makedirs(tmp_path / name)
        _json.dump(part, open(tmp_path / name / "img2txts.json", "w"))
    e = [str(tmp_path / n) for n in ("e0", "e1", "e2")]
    m = hn.get_img_txt_mappings(e)
    assert list(m[0].items()) == [("d", ["5"]), ("b", ["2"]), ("c", ["3", "4"]), ("a", ["0", "1"])]
    assert m[2] == {"d": e[2], "b": e[0], "c": e[1], "a": e[0]} and "9" not in m[1]
    assert dict(m[4]) == {e[2]: ["d"], e[0]: ["b", "a"], e[1]: ["c"]} and dict(m[5]) == {e[2]: ["5"], e[0]: ["2", "0", "1"], e[1]: ["3", "4"]}
    negs = hn.random_hard_neg({"b.npz": "b.npz", "c.npz": "c.npz"}, 1, i2s, s2i)   # (a one-image set would never end)
    assert negs["b.npz"] == ["c.npz"] and negs["c.npz"] == ["b.npz"]


def test_options_surface_matches_reference(golden_dir, tmp_path, monkeypatch):
    """lightningdot_b200/options.py against the namespaces the reference's own dvl/options.py parser produced
    (tests/golden/options_surface.json, minted by oracle/make_golden.py): every flag, default, type and choice, the
    --config JSON semantics (file values apply unless the flag is on the command line) and map_db_dirs."""
    import argparse
    import json as _json
    import sys as _sys
    from lightningdot_b200 import options as opt
    gold = _json.load(open(os.path.join(golden_dir, "options_surface.json")))

    def parser():
        p = argparse.ArgumentParser()
        opt.default_params(p)
        opt.add_itm_params(p)
        opt.add_logging_params(p)
        opt.add_kd_params(p)
        return p
    monkeypatch.setattr(_sys, "argv", ["prog"])
    assert vars(opt.parse_with_config(parser(), [])) == gold["empty"]
    for name in ("flickr30k_eval_config.json", "flickr30k_ft_config.json", "coco_ft_config.json", "coco_eval_config.json"):
        want = dict(gold[name])
        # the config file as the reference ships it, up to keys that equal the defaults: rebuilt from the parsed result
        cfg = {k: v for k, v in want.items() if k != "config" and (k not in gold["empty"] or gold["empty"][k] != v)}
        path = tmp_path / name
        _json.dump(cfg, open(path, "w"))
        got = vars(opt.parse_with_config(parser(), ["--config", str(path)]))
        got["config"] = name
        assert got == want, name
    # command-line flags win over the file
    cfg = {"seed": 42, "num_bb": 36, "project_dim": 768, "fp16": True}
    path = tmp_path / "o.json"
    _json.dump(cfg, open(path, "w"))
    monkeypatch.setattr(_sys, "argv", ["prog", "--seed=7", "--num_bb", "50"])
    got = opt.parse_with_config(parser(), ["--config", str(path), "--seed=7", "--num_bb", "50"])
    assert (got.seed, got.num_bb, got.project_dim, got.fp16) == (7, 50, 768, True)
    assert (gold["override"]["seed"], gold["override"]["num_bb"]) == (7, 50)
    with pytest.raises(SystemExit):
        parser().parse_args(["--retrieval_mode", "nonsense"])
    # map_db_dirs
    a = types.SimpleNamespace(pretrain_mapping="/mnt/pre", txt_db_mapping="/mnt/db", img_db_mapping=None,
                              val_txt_db="/db/val.db", val_img_db="/img/flickr", teacher_checkpoint="/pretrain/x.pt",
                              seed=3, train_img_dbs=["/img/a", "/img/b"], train_txt_dbs=["/db/a", "/other/b"])
    opt.map_db_dirs(a)
    assert vars(a) == gold["map_db_dirs"]
    # set_seed / setup_args_gpu on a CPU-only host
    ns = types.SimpleNamespace(seed=5, n_gpu=0, local_rank=-1, no_cuda=True, fp16=False)
    opt.set_seed(ns)
    opt.setup_args_gpu(ns)
    assert ns.device.type == "cpu" and ns.distributed_world_size == int(os.environ.get("WORLD_SIZE", "1"))


def test_amp_standin_scaler_logic():
    """lightningdot_b200.amp (stand-in for the apex.amp calls of train_itm.py:252-258): dynamic scale bookkeeping, the
    identity path for optimisers that do not run fp16 towers, master_params, and the skipped step after an overflow."""
    from lightningdot_b200 import amp
    sc = amp.LossScaler(enabled=True)
    assert sc.scale == amp.INIT_SCALE
    sc.update(True)
    assert sc.scale == amp.INIT_SCALE / 2 and sc.skipped == 1
    for _ in range(amp.GROWTH_INTERVAL):
        sc.update(False)
    assert sc.scale == amp.INIT_SCALE
    off = amp.LossScaler(enabled=False)
    off.update(True)
    assert off.scale == 1.0 and off.skipped == 0
    # plain torch optimiser on the CPU: no fp16 towers -> scale_loss is the identity, nothing is skipped
    lin = torch.nn.Linear(4, 2)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    assert [id(p) for p in amp.master_params(opt)] == [id(p) for p in lin.parameters()]
    model, opt2 = amp.initialize(lin, opt, opt_level="O1")
    assert model is lin and opt2 is opt and amp.initialize(lin) is lin
    loss = lin(torch.ones(3, 4)).sum()
    with amp.scale_loss(loss, opt) as scaled:
        assert scaled is loss
        scaled.backward()
    before = lin.weight.detach().clone()
    opt.step()
    assert not torch.equal(lin.weight.detach(), before)
    # forced fp16 bookkeeping on the CPU: an overflowing gradient zeroes the grads, halves the scale, skips ONE step
    opt.shadow_dtype = torch.float16
    opt._amp_scaler = None
    opt.zero_grad()
    loss = lin(torch.ones(3, 4)).sum() * float("inf")
    with amp.scale_loss(loss, opt) as scaled:
        scaled.backward()
    assert opt._amp_scaler.skipped == 1 and opt._amp_scaler.scale == amp.INIT_SCALE / 2
    assert all(not p.grad.any() for p in lin.parameters())
    before = lin.weight.detach().clone()
    opt.step()                                   # skipped
    assert torch.equal(lin.weight.detach(), before)
    loss = lin(torch.ones(3, 4)).sum()
    with amp.scale_loss(loss, opt) as scaled:
        scaled.backward()
    g = lin.weight.grad.clone()
    assert torch.allclose(g, torch.full_like(g, 3.0))     # unscaled in place: d/dW of sum over 3 rows of ones
    opt.step()
    assert not torch.equal(lin.weight.detach(), before)


def test_load_biencoder_checkpoint_variants(tmp_path):
    """dvl/models/bi_encoder.py:737-752 on the mirror: fine-tune checkpoints ({'model_dict': ...}), pre-training
    checkpoints (every key prefixed with 'bert.', extra heads dropped, strict load), and the 'no checkpoint' spellings."""
    from lightningdot_b200.bi_encoder import BiEncoder, TowerConfig, load_biencoder_checkpoint
    args = types.SimpleNamespace(img_model_type='uniter-base', img_model_config=TowerConfig(num_hidden_layers=1),
                                 img_checkpoint=None, txt_model_type='bert-base',
                                 txt_model_config=TowerConfig(num_hidden_layers=1), txt_checkpoint=None)
    torch.manual_seed(0)
    src = BiEncoder(args, project_dim=768)
    sd = {k: v.clone() for k, v in src.state_dict().items()}

    def fresh():
        torch.manual_seed(1)
        return BiEncoder(args, project_dim=768)

    def same(model):
        return all(torch.equal(v, sd[k]) for k, v in model.state_dict().items())
    # fine-tune checkpoint as trainer._save_checkpoint writes it
    p1 = str(tmp_path / "ft.pt")
    torch.save({"model_dict": sd, "optimizer_dict": {}, "scheduler_dict": {}, "offset": 0, "epoch": 1, "encoder_params": None}, p1)
    m = fresh()
    assert not same(m)
    load_biencoder_checkpoint(m, p1)
    assert same(m)
    # pre-training checkpoint: 'bert.' + key, plus heads the bi-encoder does not have
    p2 = str(tmp_path / "pre.pt")
    pre = {"bert." + k: v for k, v in sd.items()}
    pre["cls.predictions.bias"] = torch.zeros(3)
    pre["itm_output.weight"] = torch.zeros(2, 768)
    torch.save(pre, p2)
    m = fresh()
    load_biencoder_checkpoint(m, p2)
    assert same(m)
    # a pre-training checkpoint that lacks a tensor fails loudly (strict load)
    broken = dict(pre)
    broken.pop("bert.txt_model.encode_proj.3.bias")
    p3 = str(tmp_path / "broken.pt")
    torch.save(broken, p3)
    with pytest.raises(RuntimeError):
        load_biencoder_checkpoint(fresh(), p3)
    for none in (None, "", "none", "None"):
        m = fresh()
        before = {k: v.clone() for k, v in m.state_dict().items()}
        load_biencoder_checkpoint(m, none)
        assert all(torch.equal(v, before[k]) for k, v in m.state_dict().items())


def test_nll_positive_indices_are_checked_on_the_host():
    """BiEncoderNllLoss.calc (bi_encoder.py:615-656): the reference's Python list of positives is range-checked without a
    device round trip (no `.item()` in the training loop), and a wrong count is reported before any kernel runs."""
    from lightningdot_b200.bi_encoder import BiEncoderNllLoss
    q, c = torch.zeros(2, 8), torch.zeros(3, 8)
    with pytest.raises(IndexError):
        BiEncoderNllLoss().calc(q, c, None, [0, 3])
    with pytest.raises(IndexError):
        BiEncoderNllLoss().calc(q, c, None, [-1, 0])
    with pytest.raises(ValueError):
        BiEncoderNllLoss().calc(q, c, None, [0])
    with pytest.raises(IndexError):
        BiEncoderNllLoss().calc(q, c, None, torch.tensor([0, 7]))


def test_graphed_step_batch_copy_rules():
    """training._copy_leaves / _map_leaves (the static-batch plumbing of GraphedTrainStep): tensors are copied leaf by leaf
    into the captured buffers, Python lists of indices are accepted where a tensor was captured, non-tensor leaves are
    left alone, and a shape change is refused with a message that names the leaf."""
    from lightningdot_b200.training import _copy_leaves, _map_leaves
    static = _map_leaves({"txts": {"input_ids": torch.zeros(2, 3, dtype=torch.long), "note": "x"},
                          "pos_ctx_indices": torch.zeros(2, dtype=torch.long), "caps": {"input_ids": None}}, lambda t: t.clone())
    _copy_leaves(static, {"txts": {"input_ids": torch.arange(6).view(2, 3), "note": "y"}, "pos_ctx_indices": [1, 0],
                          "caps": {"input_ids": None}})
    assert static["txts"]["input_ids"].tolist() == [[0, 1, 2], [3, 4, 5]] and static["txts"]["note"] == "x"
    assert static["pos_ctx_indices"].tolist() == [1, 0]
    with pytest.raises(ValueError, match=r"batch\['txts'\]\['input_ids'\]"):
        _copy_leaves(static, {"txts": {"input_ids": torch.zeros(2, 4, dtype=torch.long), "note": "x"},
                              "pos_ctx_indices": [0, 1], "caps": {"input_ids": None}})


def test_search_knn_id_lists_helper_equals_the_reference_comprehension():
    """DenseFlatIndexer._format_result (faiss_indexers.py:85-87): the CPython helper (hostext/pylists.c) and the numpy
    formulation both reproduce the reference's nested list comprehension, including label -1 -> LAST id, and an
    out-of-range label raises IndexError as list indexing does."""
    from lightningdot_b200 import indexer as ix
    ix_obj = ix.DenseFlatIndexer.__new__(ix.DenseFlatIndexer)
    ix.DenseIndexer.__init__(ix_obj, buffer_size=10)
    n, nq, k = 5000, 300, 100
    ix_obj.index_id_to_db_id = [f"img_{i:07d}.npz" for i in range(n)]
    rng = np.random.default_rng(3)
    idx = rng.integers(0, n, size=(nq, k), dtype=np.int64)
    idx[7, 3] = -1
    scores = rng.standard_normal((nq, k)).astype(np.float32)
    want = [[ix_obj.index_id_to_db_id[i] for i in row] for row in idx]        # the reference's own formulation
    saved = list(ix._PYHOST)
    try:
        ix._PYHOST[:] = [False, None]
        helper = ix._pyhost_gather()
        if helper is None:
            pytest.skip("host extension not built (python -m lightningdot_b200.build)")
        got = ix_obj._format_result(scores, idx)
        assert [r[0] for r in got] == want and all((r[1] == scores[i]).all() for i, r in enumerate(got))
        assert isinstance(got[0][0], list)
        bad = idx.copy()
        bad[0, 0] = n
        with pytest.raises(IndexError):
            ix_obj._format_result(scores, bad)
        ix._PYHOST[:] = [True, None]                                         # numpy formulation
        got2 = ix_obj._format_result(scores, idx)
        assert [r[0] for r in got2] == want
    finally:
        ix._PYHOST[:] = saved
