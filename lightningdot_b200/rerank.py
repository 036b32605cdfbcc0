"""Host-side mirror of the retrieval stage of rerank.py: rerank.py:149-214.

The reference script first indexes a "validation" database with eval_model_on_dataloader(..., no_eval=True)
(rerank.py:149-157), then walks the TEST database in 400-query batches: encode both towers, search the image index with
the text embeddings and the text index with the image embeddings (k = 100), and accumulate Recall@{1,5,10,20,50,100} in
both directions plus the per-query rankings and every sample's input features (the material its cross-encoder
re-ranker consumes afterwards - that UNITER ITM re-ranker itself is outside the bi-encoder path).

Here a 400-query batch is one pass of the 128-query-tile regime of the fused score + top-k kernel per index; the
embeddings go from the towers to the search on the device (the reference copies them to numpy for faiss).
"""
import collections

import torch

RECALL_TOPS = (1, 5, 10, 20, 50, 100)


def build_retrieval_indexes(bi_encoder, dataloader, args, img2txt):
    """rerank.py:149-157 -> (indexer_img, indexer_txt) over every image / caption of the loader."""
    from .trainer import eval_model_on_dataloader
    _, _, indexers, _, _ = eval_model_on_dataloader(bi_encoder, dataloader, args, img2txt=img2txt, no_eval=True)
    return indexers


def retrieval_loop(bi_encoder, indexer_img, indexer_txt, dataloader_test, img2txt, tops=RECALL_TOPS, keep_features=True):
    """rerank.py:168-214 -> dict(recall_img, recall_txt, ranking_res_img, ranking_res_txt, feats_dict, total_len).

    recall_img[t]: fraction of test captions whose image is among the t best images (text -> image);
    recall_txt[t]: fraction of test SAMPLES whose image query finds one of its captions among the t best captions (the
    reference counts per sample, not per distinct image, and divides both by the number of samples, rerank.py:194-214)."""
    txt2img = {t: img for img, txts in img2txt.items() for t in txts}
    k = max(tops)
    hits_img, hits_txt = {t: 0 for t in tops}, {t: 0 for t in tops}
    ranking_res_img, ranking_res_txt = {}, {}
    feats_dict = {'imgs': collections.defaultdict(dict), 'txts': collections.defaultdict(dict)}
    total = 0
    bi_encoder.eval()
    for batch in dataloader_test:
        names = {'txts': batch['txt_index'], 'imgs': batch['img_fname']}
        if keep_features:   # every sample's inputs, keyed by id (rerank.py:173-184): per-row tensors, shared [1, L] rows
            for side in ('imgs', 'txts'):
                for key, val in batch[side].items():
                    for row, name in enumerate(names[side]):
                        if val is None:
                            feats_dict[side][name][key] = None
                        else:
                            feats_dict[side][name][key] = val[row] if val.shape[0] > row and val.shape[0] != 1 else val[0]
        with torch.no_grad():
            txt_vec, img_vec, _ = bi_encoder({k_: v for k_, v in batch.items() if k_ != 'caps'})
        res_img = [r[0] for r in indexer_img.search_knn(txt_vec, k)]
        res_txt = [r[0] for r in indexer_txt.search_knn(img_vec, k)]
        total += len(res_img)
        for ranked, txt_id in zip(res_img, batch['txt_index']):
            ranking_res_img[txt_id] = ranked
            for t in tops:
                hits_img[t] += txt2img[txt_id] in ranked[:t]
        for ranked, img_id in zip(res_txt, batch['img_fname']):
            ranking_res_txt[img_id] = ranked
            mine = set(img2txt[img_id])
            for t in tops:
                hits_txt[t] += any(c in mine for c in ranked[:t])
    n = max(total, 1)
    return dict(recall_img={t: hits_img[t] / n for t in tops}, recall_txt={t: hits_txt[t] / n for t in tops},
                ranking_res_img=ranking_res_img, ranking_res_txt=ranking_res_txt, feats_dict=feats_dict, total_len=total)
