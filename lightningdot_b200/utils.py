"""Host-side mirror of the slice of dvl/utils.py the retrieval path uses.

  _calc_loss        dvl/utils.py:114-169   in-batch-negative loss wrapper (world size 1 in the reference: its
                                           cross-rank branch is dead code, :121; here the gather is live when a
                                           torch.distributed group is initialised and args.distributed_world_size > 1)
  gather_embeddings / sync_gradients       the two exchange steps of global-batch training: a
                                           differentiable all-gather of the embeddings and the gradient average
  retrieve_query    dvl/utils.py:204-211   free-text query -> txt tower -> search_knn(., 100)
  get_model_encoded_vecs   dvl/utils.py:214-234   encode a loader -> {id: embedding} dictionaries (the demo's index source)
  all_gather_list   dvl/utils.py:51-111    arbitrary picklable data from every rank
  is_main_process / get_rank / get_world_size   dvl/utils.py:18-23,187-188 (horovod there; env / torch.distributed here)
  print_args, num_of_parameters, compare_models dvl/utils.py:26-38,172-184
"""
import logging
import os
import pickle
from collections import defaultdict

import torch
import torch.distributed as dist

logger = logging.getLogger()


def get_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return int(os.environ.get("RANK", "0"))


def get_world_size():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return int(os.environ.get("WORLD_SIZE", "1"))


def is_main_process():
    return get_rank() == 0


def print_args(args):
    logger.info(" **************** CONFIGURATION **************** ")
    for key, val in sorted(vars(args).items()):
        logger.info("%s -->   %s", f"{key:<30}", val)
    logger.info(" **************** CONFIGURATION **************** ")


def num_of_parameters(model, requires_grad=False):
    return sum(p.numel() for p in model.parameters() if p.requires_grad or not requires_grad)


def compare_models(model_1, model_2):
    differ = [k1 for (k1, v1), (k2, v2) in zip(model_1.state_dict().items(), model_2.state_dict().items())
              if k1 != k2 or not torch.equal(v1, v2)]
    for k in differ:
        print('Mismtach found at', k)
    if not differ:
        print('Models match perfectly! :)')
    return len(differ)


def _mean_or_sum_(t, group, mean):
    """In-place all-reduce (NCCL has AVG; gloo only SUM)."""
    if mean and dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
        return t
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    if mean:
        t.div_(dist.get_world_size(group))
    return t


class _GatherWithGrad(torch.autograd.Function):
    """[b, D] on every rank -> [W * b, D] in rank order, differentiable.  Every rank's loss reads ALL gathered rows, so
    the gradient of the SUM of the ranks' losses w.r.t. this rank's block is the reduce-scatter (sum) of the ranks'
    d(gathered) - one collective in backward, called by every rank in the same order."""

    @staticmethod
    def forward(ctx, local, group):
        world = dist.get_world_size(group)
        ctx.group, ctx.rank, ctx.rows = group, dist.get_rank(group), local.shape[0]
        src = local.contiguous()
        out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        dist.all_gather(list(out.chunk(world, dim=0)), src, group=group)
        return out

    @staticmethod
    def backward(ctx, d_out):
        d_out = d_out.contiguous()
        lo = ctx.rank * ctx.rows
        if dist.get_backend(ctx.group) == "nccl":
            d_local = torch.empty_like(d_out[lo:lo + ctx.rows])
            dist.reduce_scatter_tensor(d_local, d_out, op=dist.ReduceOp.SUM, group=ctx.group)
        else:   # gloo has no reduce-scatter
            d_local = _mean_or_sum_(d_out.clone(), ctx.group, mean=False)[lo:lo + ctx.rows].clone()
        return d_local, None


def gather_embeddings(local, group=None):
    """All-gather a [b, D] embedding block over the ranks (equal b on every rank) -> [W * b, D].  The reference's
    sketch (dvl/utils.py:143-152) keeps grad on the local slice only, which drops the terms other ranks' queries
    contribute to this rank's contexts; here the gather is differentiable (backward = reduce-scatter), so with
    gradients AVERAGED over ranks (sync_gradients / FusedAdamW.distributed) a W-rank step equals the single-process
    step on the global batch."""
    if torch.is_grad_enabled() and local.requires_grad:
        return _GatherWithGrad.apply(local, group)
    world = dist.get_world_size(group)
    src = local.detach().contiguous()
    out = torch.empty((world * src.shape[0],) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    dist.all_gather(list(out.chunk(world, dim=0)), src, group=group)
    return out


def sync_gradients(params_or_buffers, group=None):
    """Average gradients over the ranks, in place (DDP's convention: every rank's loss is the mean over ITS query
    rows, so the mean over ranks is the global-batch mean).  Accepts parameters (their .grad is reduced) or plain
    tensors (FusedAdamW passes its flat gradient buffers: one collective per parameter group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in params_or_buffers:
        g = t.grad if isinstance(t, torch.nn.Parameter) else t
        if g is not None:
            _mean_or_sum_(g, group, mean=True)


def _calc_loss(args, loss_function, local_q_vector, local_ctx_vectors, local_caption_vectors, local_positive_idxs,
               local_hard_negatives_idxs: list = None, experiment=None):
    """In-batch-negatives loss (dvl/utils.py:114-169).  With one process this is loss_function.calc on the local
    batch; with a process group and args.distributed_world_size > 1 the embeddings of all ranks are gathered and
    positives are shifted by the rank's context offset, which is what the reference's (disabled) branch sketches."""
    world = int(getattr(args, "distributed_world_size", 1) or 1)
    if world > 1 and dist.is_available() and dist.is_initialized():
        rank = dist.get_rank()
        n_ctx = local_ctx_vectors.shape[0]
        q = local_q_vector
        ctx = gather_embeddings(local_ctx_vectors)
        cap = gather_embeddings(local_caption_vectors) if local_caption_vectors is not None else None
        if torch.is_tensor(local_positive_idxs):
            positives = local_positive_idxs + rank * n_ctx
        else:
            positives = [p + rank * n_ctx for p in local_positive_idxs]
        hard = None if local_hard_negatives_idxs is None else \
            [[v + rank * n_ctx for v in row] for row in local_hard_negatives_idxs]
    else:
        q, ctx, cap = local_q_vector, local_ctx_vectors, local_caption_vectors
        positives, hard = local_positive_idxs, local_hard_negatives_idxs
    return loss_function.calc(q, ctx, cap, positives, hard, getattr(args, "caption_score_weight", 0.0), experiment)


def retrieve_query(model, query, indexer, args, top=10):
    """dvl/utils.py:204-211 (the reference ignores `top` and always asks for 100; kept)."""
    input_ids = torch.as_tensor(args.tokenizer.encode(query), dtype=torch.long, device=args.device).unsqueeze(0)
    n = input_ids.shape[1]
    attn_mask = torch.ones((1, n), dtype=torch.long, device=args.device)
    pos_ids = torch.arange(n, dtype=torch.long, device=args.device).unsqueeze(0)
    with torch.no_grad():
        _, query_vector, _ = model.txt_model(input_ids=input_ids, attention_mask=attn_mask, position_ids=pos_ids,
                                             need_sequence=False)
    return indexer.search_knn(query_vector, 100)


def get_model_encoded_vecs(model, dataloader):
    """dvl/utils.py:214-234 -> {'img_embed': {img id: vec}, 'caption_embed': {img id: vec}, 'txt_embed': {txt id: vec},
    'img_name': ids of the LAST batch (the reference extends the list after the loop, :227)}.  One device -> host copy
    per batch and tower instead of one per row; later encodings of an id overwrite earlier ones, as dict.update does."""
    img_embedding, caption_embedding, query_embedding = dict(), dict(), defaultdict(list)
    batch = None
    for batch in dataloader:
        with torch.no_grad():
            q_vec, ctx_vec, cap_vec = model(batch)
        for store, keys, vecs in ((img_embedding, batch['img_fname'], ctx_vec),
                                  (caption_embedding, batch['img_fname'], cap_vec),
                                  (query_embedding, batch['txt_index'], q_vec)):
            if vecs is None:   # (no caption tower input: the reference would fail on zip(None); nothing to record)
                continue
            host = vecs.detach().float().cpu().numpy()
            store.update({key: host[row] for row, key in enumerate(keys)})
    return {'img_embed': img_embedding, 'caption_embed': caption_embedding, 'txt_embed': query_embedding,
            'img_name': list(batch['img_fname']) if batch is not None else []}


def all_gather_list(data, group=None, max_size=16384):
    """dvl/utils.py:51-111: gather arbitrary picklable `data` from every rank -> list in rank order.  Same contract
    (ValueError when the pickle plus its 4-byte length prefix exceeds max_size; every rank must call it), carried by one
    all_gather of fixed-size byte rows (the reference all-reduces a zero-padded [world * max_size] CUDA byte buffer)."""
    enc = pickle.dumps(data)
    if len(enc) + 4 > max_size:
        raise ValueError(f'encoded data exceeds max_size, this can be fixed by increasing buffer size: {len(enc)}')
    if not (dist.is_available() and dist.is_initialized()):
        return [pickle.loads(enc)]
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    row = torch.zeros(max_size, dtype=torch.uint8)
    row[:4] = torch.tensor(list(len(enc).to_bytes(4, byteorder='big')), dtype=torch.uint8)
    row[4:4 + len(enc)] = torch.frombuffer(bytearray(enc), dtype=torch.uint8)
    row = row.to(dev)
    rows = torch.empty((world, max_size), dtype=torch.uint8, device=dev)
    dist.all_gather(list(rows.unbind(0)), row, group=group)
    rows = rows.cpu()
    out = []
    for r in range(world):
        n = int.from_bytes(bytes(rows[r, :4].tolist()), byteorder='big')
        if n > 0:
            out.append(pickle.loads(bytes(rows[r, 4:4 + n].numpy().tobytes())))
    return out
