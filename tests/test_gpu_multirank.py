"""Multi-GPU paths on real devices: one process per GPU over NCCL.  Needs >= 2 GPUs on the box
(`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`); skipped on a single-GPU box, where the same
protocols are covered over gloo by tests/test_host_logic.py.

  e1  row-sharded exact search == the single-GPU search, bit for bit (ids and scores), every rank
  e1/e2 through the reference's entry points: eval_itm.py's flow under a 2-rank group (captions strided over the ranks,
      embeddings pooled, ShardedFlatIndexer from _new_indexer) == the single-process fixture
  e3  global-batch in-batch NLL: the 2-rank training step (differentiable embedding all-gather + gradient average)
      == the same global batch on one GPU (loss and every parameter gradient)
"""
import json
import os
import socket
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run under gpurun --gpus 2)")]


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _init(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), LOCAL_WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    import sys
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)


def _search_worker(rank, world, port, tmp):
    _init(rank, world, port)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    try:
        from lightningdot_b200 import synth
        from lightningdot_b200.indexer import DenseFlatIndexer
        from lightningdot_b200.sharded import ShardedFlatIndexer
        n, nq, k = 60001, 333, 100
        x = synth.gaussian_index(n, 768, seed=21)
        q, _ = synth.planted_queries(x, nq, sigma=2.0, seed=22)
        ids = [f"img_{i:07d}.npz" for i in range(n)]
        sh = ShardedFlatIndexer(768)
        sh.index_matrix(ids, torch.from_numpy(x).cuda())
        assert sh.index.ntotal == sh.bounds[rank + 1] - sh.bounds[rank]
        qd = torch.from_numpy(q).cuda()
        s_sh, i_sh = sh.search_device(qd, k)
        one = DenseFlatIndexer(768)
        one.index_matrix(ids, torch.from_numpy(x).cuda())
        s_1, i_1 = one.index.search_device(qd, k)
        assert torch.equal(i_sh, i_1) and torch.equal(s_sh, s_1)
        for nq2, k2 in ((5, 7), (1, 1), (64, 33)):     # odd slice sizes (padding of the packed exchange), fewer queries than ranks
            a_s, a_i = sh.search_device(qd[:nq2].contiguous(), k2)
            b_s, b_i = one.index.search_device(qd[:nq2].contiguous(), k2)
            assert torch.equal(a_i, b_i) and torch.equal(a_s, b_s), (nq2, k2)
        l_s, l_i = sh.search_device(qd, k, lazy_flags=True)
        assert sh.pending_flags() == 0 and torch.equal(l_i, i_1) and torch.equal(l_s, s_1)
        assert (sh.search(q[:9], 10, host_rank=0) is None) == (rank != 0)
        res = sh.search_knn(q[:7], 10)
        ref = one.search_knn(q[:7], 10)
        assert [r[0] for r in res] == [r[0] for r in ref]
        # queries encoded nq / W per rank and gathered
        b = [0, nq // 2, nq]
        assert torch.equal(sh.gather_queries(qd[b[rank]:b[rank + 1]]), qd)
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_row_sharded_search_equals_single_gpu_search(tmp_path):
    mp.spawn(_search_worker, args=(2, _port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))


def _eval_flow_worker(rank, world, port, tmp, gold_path):
    _init(rank, world, port)
    try:
        import itm_flow_tool as flow
        gold = json.load(open(gold_path))
        ws_dir = os.path.join(tmp, "ws")
        if rank == 0:
            flow.make_workspace(ws_dir, layers=gold["layers"], seed_txt=gold["seed_txt"], seed_img=gold["seed_img"],
                                batch_size=gold["batch_size"], seed_db=gold["db"]["seed"])
            open(os.path.join(tmp, "ready"), "w").write("1")
        import time
        while not os.path.exists(os.path.join(tmp, "ready")):
            time.sleep(0.1)
        out = flow.eval_flow(os.path.join(ws_dir, "eval_config.json"), os.path.join(ws_dir, "ckpt_run", "biencoder.last.pt"),
                             fp16=True)            # hvd.init() inside joins the 2-rank NCCL group
        from lightningdot_b200.sharded import ShardedFlatIndexer
        ix_img, ix_txt = out["indexers"]
        assert isinstance(ix_img, ShardedFlatIndexer) and ix_img.n_global == 40 and ix_txt.n_global == 200
        assert ix_img.index.ntotal == 20 and ix_txt.index.ntotal == 100
        assert len(out["rank_txt"]) == 200 and len(out["rank_img"]) == 40
        for key, one in (("recall_txt", 1 / 200), ("recall_img", 1 / 40)):
            for t in (1, 5, 10):
                assert abs(out[key][t] - gold[key][str(t)]) <= one + 1e-9, (key, t, out[key], gold[key])
        same = np.mean([len(set(out["rank_txt"][q][:10]) & set(v)) / 10 for q, v in gold["rank_txt_top10"].items()])
        assert same >= 0.97, same
        json.dump({"recall_txt": out["recall_txt"], "recall_img": out["recall_img"], "loss": out["loss"],
                   "top": {q: out["rank_txt"][q][:10] for q in ("0", "7", "199")}},
                  open(os.path.join(tmp, f"res{rank}.json"), "w"))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def test_eval_flow_two_ranks_through_namesake_packages(tmp_path, golden_dir):
    mp.spawn(_eval_flow_worker, args=(2, _port(), str(tmp_path), os.path.join(golden_dir, "evalflow_small.json")),
             nprocs=2, join=True)
    r0, r1 = (json.load(open(tmp_path / f"res{r}.json")) for r in range(2))
    assert r0 == r1     # every rank holds the same merged result


def _train_worker(rank, world, port, tmp):
    _init(rank, world, port)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from lightningdot_b200 import synth
        from lightningdot_b200.bi_encoder import (BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer,
                                                  setup_for_distributed_mode)
        from lightningdot_b200.utils import _calc_loss
        b, layers = 24, 2
        B = b * world
        torch.manual_seed(5)
        cfg = dict(img_model_type='uniter-base', img_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers),
                   img_checkpoint=None, txt_model_type='bert-base',
                   txt_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers), txt_checkpoint=None)
        model = BiEncoder(types.SimpleNamespace(**cfg), project_dim=768)
        opt = get_optimizer(model, learning_rate=1e-5)
        model, opt = setup_for_distributed_mode(model, opt, dev, 1, rank, True)     # fp16 towers; broadcast from rank 0
        model.eval()          # dropout off: the two runs would otherwise draw different masks; gradients still flow
        tb, ib = synth.text_batch(B, 32, seed=7, ragged=True), synth.image_batch(B, 36, seed=8, ragged=True)

        def batch(lo, hi):
            def sl(d):
                return {k: (v[lo:hi].contiguous().to(dev) if (torch.is_tensor(v) and v.shape[0] == B) else
                            (v.to(dev) if torch.is_tensor(v) else v)) for k, v in d.items()}
            return {"txts": sl(tb), "imgs": sl(ib), "caps": {"input_ids": None}, "pos_ctx_indices": list(range(hi - lo))}

        def run(bt, la):
            t, i, _ = model(bt)
            l1, _, _ = _calc_loss(la, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
            l2, _, _ = _calc_loss(la, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
            loss = 0.5 * l1 + 0.5 * l2
            loss.backward()
            return loss.detach()

        loss_d = run(batch(rank * b, rank * b + b), types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=world))
        opt.sync_gradients()
        g_dist = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        dist.all_reduce(loss_d, op=dist.ReduceOp.AVG)
        model.zero_grad()
        loss_s = run(batch(0, B), types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=1))
        assert abs(loss_d.item() - loss_s.item()) <= 2e-4 * abs(loss_s.item()), (loss_d.item(), loss_s.item())
        num = den = 0.0
        for n, p in model.named_parameters():
            if p.grad is not None:
                num += float((g_dist[n] - p.grad).norm()) ** 2
                den += float(p.grad.norm()) ** 2
        rel = (num / den) ** 0.5
        assert rel <= 5e-3, rel     # 16-bit activation gradients take different reduction orders in the two runs
        # overlapped gradient averaging (FusedAdamW.overlap_grad_sync): spans of the flat buffers are all-reduced while the
        # backward is still running; the result must be the gradient average of the plain path
        opt._collect()
        opt.zero_grad()       # (also forgets that the gradients above were synchronised)
        opt.overlap_grad_sync, opt.early_sync_bytes = True, 4 << 20
        la = types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=world)
        run(batch(rank * b, rank * b + b), la)
        torch.cuda.synchronize()
        early = len(opt._early_spans)
        covered = sum(hi - lo for _, lo, hi in opt._early_spans)
        opt.sync_gradients()
        total = sum(f["g"].numel() for f in opt._flat if f is not None)
        assert early >= 6 and covered >= 0.5 * total, (early, covered, total)
        assert not opt._early_spans and not opt._early_works
        num = den = 0.0
        for n, p in model.named_parameters():
            if p.grad is not None:
                num += float((g_dist[n] - p.grad).norm()) ** 2
                den += float(g_dist[n].norm()) ** 2
        rel2 = (num / den) ** 0.5
        assert rel2 <= 1e-4, rel2   # (same computation; wgrad / column-sum atomics reorder)
        open(os.path.join(tmp, f"ok{rank}"), "w").write(f"{loss_s.item()} {rel} {rel2}")
    finally:
        dist.destroy_process_group()


def test_global_batch_training_step_equals_single_gpu_step(tmp_path):
    mp.spawn(_train_worker, args=(2, _port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))


def _graph_worker(rank, world, port, tmp):
    _init(rank, world, port)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    gstep = None
    try:
        from lightningdot_b200 import synth
        from lightningdot_b200.bi_encoder import (BiEncoder, BiEncoderNllLoss, TowerConfig, get_optimizer, get_schedule_linear,
                                                  setup_for_distributed_mode)
        from lightningdot_b200.training import GraphedTrainStep
        from lightningdot_b200.utils import _calc_loss
        b, layers, steps = 16, 2, 5
        B = b * world
        torch.manual_seed(9)
        cfg = dict(img_model_type='uniter-base', img_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers),
                   img_checkpoint=None, txt_model_type='bert-base',
                   txt_model_config=TowerConfig(vocab_size=synth.VOCAB, num_hidden_layers=layers), txt_checkpoint=None)
        model = BiEncoder(types.SimpleNamespace(**cfg), project_dim=768)
        opt = get_optimizer(model, learning_rate=2e-6, adam_eps=1e-4, weight_decay=0.01)
        opt.max_grad_norm = 2.0
        model, opt = setup_for_distributed_mode(model, opt, dev, 1, rank, False)
        model.train()                                   # dropout on
        opt.overlap_grad_sync, opt.early_sync_bytes = True, 4 << 20
        sched = get_schedule_linear(opt, 2, 50)
        la = types.SimpleNamespace(caption_score_weight=0.0, distributed_world_size=world)

        def batch(seed):
            tb, ib = synth.text_batch(B, 32, seed=seed, ragged=True), synth.image_batch(B, 36, seed=seed + 50, ragged=True)
            lo, hi = rank * b, rank * b + b

            def sl(d):
                return {k: (v[lo:hi].contiguous() if (torch.is_tensor(v) and v.shape[0] == B) else v) for k, v in d.items()}
            return {"txts": sl(tb), "imgs": sl(ib), "caps": {"input_ids": None}, "pos_ctx_indices": list(range(b))}

        def fwd_bwd(bt):
            t, i, _ = model(bt)
            l1, _, _ = _calc_loss(la, BiEncoderNllLoss(), i, t, None, bt["pos_ctx_indices"], None)
            l2, _, _ = _calc_loss(la, BiEncoderNllLoss(), t, i, None, bt["pos_ctx_indices"], None)
            loss = 0.5 * l1 + 0.5 * l2
            loss.backward()
            return loss

        gstep = GraphedTrainStep(fwd_bwd, opt, batch(0), scheduler=sched, warmup=1)
        losses = [gstep(batch(s)).item() for s in range(2, steps)]
        taken, epoch = gstep.steps_taken, int(gstep.epoch.item())
        gstep.release()
        gstep = None
        assert all(np.isfinite(losses)), losses
        assert taken == steps and epoch == steps - 2    # (two eager warm-up steps: layout + steady state)
        # the ranks took the same steps: parameters stay identical across the group (gradients were averaged inside the
        # captured step - embedding all-gather, reduce-scatter, overlapped all-reduce are graph nodes)
        worst = 0.0
        for f in opt._flat:
            if f is not None:
                ref = f["p"].clone()
                dist.broadcast(ref, src=0)
                worst = max(worst, (ref - f["p"]).abs().max().item())
        ok = worst <= 1e-7 and max(f["m"].abs().max().item() for f in opt._flat if f is not None) > 0
        open(os.path.join(tmp, f"{'ok' if ok else 'bad'}{rank}"), "w").write(json.dumps([losses, worst]))
    finally:
        if gstep is not None:
            gstep.release()     # (a live graph that captured NCCL collectives makes destroy_process_group() block)
        dist.destroy_process_group()


def test_graphed_train_step_two_ranks(tmp_path):
    """training.GraphedTrainStep under a 2-rank NCCL group: the captured step contains the collectives of the global-batch
    loss and the overlapped gradient average; replays keep the ranks' parameters identical (to the last bits of the
    atomics-ordered clip norm)."""
    mp.spawn(_graph_worker, args=(2, _port(), str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(2))
