"""Host-side mirror of the evaluation half of dvl/trainer.py (the encode -> index -> search -> Recall@k driver).

  eval_model_on_dataloader   dvl/trainer.py:113-190   same arguments, same 5-tuple return
  get_indexer                dvl/trainer.py:93-110
  build_dataloader           dvl/trainer.py:28-37     DataLoader + PrefetchLoader over a dataset of data.py
  load_dataset               dvl/trainer.py:193-209   text / image database folders -> ItmFastDataset(s)
  CheckpointState, _save_checkpoint, load_saved_state, load_states_from_checkpoint   dvl/trainer.py:18-20,44-90

What differs from the reference is where the data lives, not what is computed: the reference copies every embedding
row to the host one `.cpu().numpy()` at a time (trainer.py:135,138,151-152) and hands numpy matrices to faiss; here the
embeddings of a batch stay in HBM, the "last encoding wins, first-seen order" de-duplication of trainer.py:151-152
(dict.update semantics) is done on row numbers, and the two indexes are built from device gathers.  Results (loss,
accuracy, recalls, ranked id lists) are the same objects the reference returns.
"""
import collections
import logging
import os

import numpy as np
import torch
import torch.distributed as dist

from torch.utils.data import ConcatDataset, DataLoader

from .bi_encoder import BiEncoderNllLoss
from .data import ItmFastDataset, ItmValDataset, TxtTokLmdb, itm_fast_collate  # noqa: F401  (reference import surface)
from .indexer import DenseFlatIndexer, DenseHNSWFlatIndexer
from .loader import PrefetchLoader
from .utils import _calc_loss

logger = logging.getLogger()

CheckpointState = collections.namedtuple(
    "CheckpointState", ['model_dict', 'optimizer_dict', 'scheduler_dict', 'offset', 'epoch', 'encoder_params'])


class BiEncoderTrainer:
    """Empty in the reference too (dvl/trainer.py:23-25); the scripts drive the functions below."""

    def __init__(self, args):
        pass


def build_dataloader(dataset, collate_fn, is_train, opts, batch_size=None):
    """dvl/trainer.py:28-37: shuffled (training) or sequential (evaluation) batches of opts.train_batch_size /
    opts.valid_batch_size samples, collated on worker processes, moved to the GPU one batch ahead on a side stream."""
    if batch_size is None:
        batch_size = opts.train_batch_size if is_train else opts.valid_batch_size
    loader = DataLoader(dataset, batch_size=batch_size, shuffle=is_train, drop_last=False, num_workers=opts.n_workers,
                        pin_memory=opts.pin_mem, collate_fn=collate_fn)
    return PrefetchLoader(loader)


def load_dataset(all_img_dbs, txt_dbs, img_dbs, args, is_train):
    """dvl/trainer.py:193-209.  Training: one ItmFastDataset per (text folder, image folder) pair, captions no longer than
    args.max_txt_len, concatenated.  Evaluation: ONE dataset over a single folder pair with every caption kept (the
    reference passes args.inf_minibatch_size in the hard-negative slot here; without mined negatives it has no effect)."""
    if is_train:
        return ConcatDataset([ItmFastDataset(TxtTokLmdb(txt_path, args.max_txt_len), all_img_dbs[img_path],
                                             args.num_hard_negatives, args.img_meta, args.tokenizer)
                              for txt_path, img_path in zip(txt_dbs, img_dbs)])
    return ItmFastDataset(TxtTokLmdb(txt_dbs, -1), all_img_dbs[img_dbs], args.inf_minibatch_size, args.img_meta,
                          args.tokenizer)


def get_model_obj(model):
    return model.module if hasattr(model, 'module') else model


def _save_checkpoint(args, biencoder, optimizer, scheduler, epoch: int, offset: int, cp_name: str = None) -> str:
    """dvl/trainer.py:44-63: biencoder.<name>.pt holding a CheckpointState dict."""
    stem = cp_name if cp_name is not None else str(epoch) + ('.' + str(offset) if offset > 0 else '')
    cp = os.path.join(args.output_dir, 'biencoder.' + stem + '.pt')
    state = CheckpointState(get_model_obj(biencoder).state_dict(), optimizer.state_dict(), scheduler.state_dict(),
                            offset, epoch, None)
    torch.save(state._asdict(), cp)
    logger.info('Saved checkpoint at %s', cp)
    return cp


def load_states_from_checkpoint(model_file: str) -> CheckpointState:
    logger.info('Reading saved model from %s', model_file)
    state_dict = torch.load(model_file, map_location='cpu')
    logger.info('model_state_dict keys %s', state_dict.keys())
    return CheckpointState(**state_dict)


def load_saved_state(biencoder, optimizer=None, scheduler=None, saved_state: CheckpointState = ''):
    """dvl/trainer.py:66-83."""
    epoch = saved_state.epoch + (1 if saved_state.offset == 0 else 0)  # offset 0: that epoch was completed
    logger.info('Loading checkpoint @ batch=%s and epoch=%s', saved_state.offset, epoch)
    get_model_obj(biencoder).load_state_dict(saved_state.model_dict)
    if saved_state.optimizer_dict and optimizer is not None:
        optimizer.load_state_dict(saved_state.optimizer_dict)
    if saved_state.scheduler_dict and scheduler is not None:
        scheduler.load_state_dict(saved_state.scheduler_dict)


class _LastWins:
    """dict.update({id: vec}) semantics on device rows: keys keep first-seen order, values keep the last row."""

    def __init__(self):
        self.row_of = {}

    def update(self, ids, first_row):
        row_of = self.row_of
        for j, key in enumerate(ids):
            row_of[key] = first_row + j

    def keys(self):
        return list(self.row_of.keys())

    def rows(self, device):
        return torch.as_tensor(list(self.row_of.values()), dtype=torch.int64, device=device)


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def _new_indexer(args, hnsw_index):
    """dvl/trainer.py:93-98,160-161.  One process: DenseFlatIndexer.  Under a torch.distributed group (one process per GPU)
    the index is row-sharded over the ranks (ShardedFlatIndexer: per-shard exact top-k, NCCL all-gather, merge) - the
    8-GPU configuration of BASELINE configs[3], reached through the reference's own entry points."""
    if hnsw_index:
        return DenseHNSWFlatIndexer(args.vector_size)
    if _world() > 1 and getattr(args, "shard_index", True):
        from .sharded import ShardedFlatIndexer
        return ShardedFlatIndexer(args.vector_size)
    return DenseFlatIndexer(args.vector_size)


def _gather_interleaved(local, ids, world):
    """Every rank encoded the samples ids[rank::world] of one sequential pass (TxtTokLmdb strides the caption ids over
    the ranks, uniter_model/data/data.py:186-187).  -> (the [n_total, D] matrix, the id lists) in the order ONE process
    would have seen them, identical on every rank: sample j of rank r is global sample j * world + r."""
    dev = local.device
    counts = [None] * world
    dist.all_gather_object(counts, (local.shape[0], list(ids[0]), list(ids[1])))
    m = max(c[0] for c in counts)
    padded = local.new_zeros((m, local.shape[1]))
    padded[:local.shape[0]] = local
    parts = torch.empty((world, m, local.shape[1]), dtype=local.dtype, device=dev)
    dist.all_gather_into_tensor(parts.view(world * m, -1), padded)
    total = sum(c[0] for c in counts)
    out = local.new_empty((total, local.shape[1]))
    ids_a, ids_b = [None] * total, [None] * total
    for r, (n, a, b) in enumerate(counts):
        pos = torch.arange(n, device=dev) * world + r
        if n and int(pos[-1]) >= total:
            raise ValueError("ranks hold unevenly strided shares of the dataset: expected ids[rank::world]")
        out[pos] = parts[r, :n]
        ids_a[r::world] = a
        ids_b[r::world] = b
    return out, ids_a, ids_b


def get_indexer(bi_encoder, eval_dataloader, args, hnsw_index, img_retrieval=True):
    """dvl/trainer.py:93-110: encode a whole loader and index the image (or text) side."""
    bi_encoder.eval()
    indexer = _new_indexer(args, hnsw_index)
    seen, chunks, n = _LastWins(), [], 0
    for batch in eval_dataloader:
        with torch.no_grad():
            q_vec, ctx_vec, _ = bi_encoder(batch)
        vec, ids = (ctx_vec, batch['img_fname']) if img_retrieval else (q_vec, batch['txt_index'])
        chunks.append(vec.detach().float())
        seen.update(ids, n)
        n += vec.shape[0]
    if n:
        allv = torch.cat(chunks, 0)
        world = _world()
        if world > 1 and getattr(args, "shard_index", True) and not hnsw_index:
            ids = [k for k, _ in sorted(seen.row_of.items(), key=lambda kv: kv[1])] if len(seen.row_of) == n else None
            if ids is None:   # duplicates inside a rank's share: resolve them locally first
                allv, ids = allv.index_select(0, seen.rows(allv.device)), seen.keys()
            allv, ids, _ = _gather_interleaved(allv, (ids, ids), world)
            seen = _LastWins()
            seen.update(ids, 0)
        indexer.index_matrix(seen.keys(), allv.index_select(0, seen.rows(allv.device)))
    return indexer


def eval_model_on_dataloader(bi_encoder, eval_dataloader, args, img2txt=None, num_tops=100, no_eval=False):
    """dvl/trainer.py:113-190 -> (loss, acc, (indexer_img, indexer_txt), (recall_txt, recall_img),
    (rank_txt_res, rank_img_res))."""
    bi_encoder.eval()
    indexer_img = _new_indexer(args, args.hnsw_index)
    indexer_txt = _new_indexer(args, args.hnsw_index)
    loss_function = BiEncoderNllLoss()
    total_loss, total_correct = None, None
    batches = total_samples = rows = 0
    txt_chunks, img_chunks = [], []
    query_txt_id, query_img_id = [], []
    img_seen, txt_seen = _LastWins(), _LastWins()

    for batch in eval_dataloader:
        with torch.no_grad():
            q_vec, ctx_vec, cap_vec = bi_encoder(batch)
            loss, correct_cnt, _ = _calc_loss(args, loss_function, q_vec, ctx_vec, cap_vec,
                                              list(range(len(q_vec))), None)
        # running sums stay on the device: one host sync at the end instead of two .item() per batch
        total_loss = loss.detach().double() if total_loss is None else total_loss + loss.detach().double()
        total_correct = correct_cnt.sum() if total_correct is None else total_correct + correct_cnt.sum()
        batches += 1
        total_samples += batch['txts']['input_ids'].shape[0]
        txt_chunks.append(q_vec.detach().float())
        img_chunks.append(ctx_vec.detach().float())
        query_txt_id.extend(batch['txt_index'])
        query_img_id.extend(batch['img_fname'])
        img_seen.update(batch['img_fname'], rows)
        txt_seen.update(batch['txt_index'], rows)
        rows += q_vec.shape[0]

    if batches == 0:
        raise ValueError("empty dataloader")
    query_txt = torch.cat(txt_chunks, 0)
    query_img = torch.cat(img_chunks, 0)
    dev = query_txt.device
    world = _world()
    if world > 1 and getattr(args, "shard_index", True):
        # every rank evaluated its stride of the captions: pool the embeddings (one all-gather per tower), rebuild the
        # single-process sample order, and let every rank index / search the GLOBAL set through the row-sharded indexer.
        # (The reference would report each rank's recall over its own subset; this is the one-process result.)
        sums = torch.stack([total_loss, total_correct.double(),
                            torch.tensor(float(batches), device=dev, dtype=torch.float64),
                            torch.tensor(float(total_samples), device=dev, dtype=torch.float64)])
        dist.all_reduce(sums)
        total_loss, total_correct = sums[0], sums[1]
        batches, total_samples = int(sums[2].item()), int(sums[3].item())
        both, query_txt_id, query_img_id = _gather_interleaved(torch.cat([query_txt, query_img], 1),
                                                               (query_txt_id, query_img_id), world)
        query_txt, query_img = both[:, :query_txt.shape[1]].contiguous(), both[:, query_txt.shape[1]:].contiguous()
        img_seen, txt_seen = _LastWins(), _LastWins()
        img_seen.update(query_img_id, 0)
        txt_seen.update(query_txt_id, 0)
    total_loss = float(total_loss.item()) / batches
    correct_ratio = float(total_correct.item()) / float(total_samples)
    indexer_img.index_matrix(img_seen.keys(), query_img.index_select(0, img_seen.rows(dev)))
    indexer_txt.index_matrix(txt_seen.keys(), query_txt.index_select(0, txt_seen.rows(dev)))
    if no_eval:
        return total_loss, correct_ratio, (indexer_img, indexer_txt), (None, None), (None, None)

    labels_img_name = query_img_id
    res_txt = indexer_img.search_knn(query_txt, num_tops)
    rank_txt_res = {query_txt_id[i]: r[0] for i, r in enumerate(res_txt)}
    res_img = indexer_txt.search_knn(query_img, num_tops)
    rank_img_res = {query_img_id[i]: r[0] for i, r in enumerate(res_img)}

    tops = (1, 5, 10)
    recall_txt = {t: 0 for t in tops}
    for i, q in enumerate(query_txt_id):
        ranked = rank_txt_res[q]
        for t in tops:
            recall_txt[t] += labels_img_name[i] in ranked[:t]
    for t in tops:
        recall_txt[t] = recall_txt[t] / len(rank_txt_res)

    recall_img = {t: 0 for t in tops}
    for q in np.unique(query_img_id):
        ranked = rank_img_res[q]
        for t in tops:
            recall_img[t] += any(txt_id in ranked[:t] for txt_id in img2txt[q])
    for t in tops:
        recall_img[t] = recall_img[t] / len(rank_img_res)

    return total_loss, correct_ratio, (indexer_img, indexer_txt), (recall_txt, recall_img), (rank_txt_res, rank_img_res)
