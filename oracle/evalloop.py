"""TEST INFRASTRUCTURE (parity oracle) - the encode -> index -> search -> Recall@1/5/10 loop, CPU.

Pinned by oracle/make_golden.py against the reference's own eval_model_on_dataloader (imported from
/root/reference, faiss replaced by the numpy stand-in of oracle/ref_shims.py) -> tests/golden/evalloop_*.json.

Follows dvl/trainer.py:113-190.  Quirks kept on purpose: every sample is a caption with
its paired image, so an image is encoded once per caption and the index keeps the LAST encoding (dict.update);
query_img keeps all duplicates while rank_img_res keeps the last; recall_img divides by the number of distinct
images.
"""
import numpy as np

from . import flatip


def recall_from_embeddings(txt_vecs, img_vecs, txt_ids, img_ids, img2txt, num_tops=100, scorer=flatip.scores_f64):
    """txt_vecs[i] / img_vecs[i]: embeddings of sample i (caption i and ITS image, re-encoded per caption);
    txt_ids[i] / img_ids[i]: their names.  -> (recall_txt, recall_img, rank_txt_res, rank_img_res)"""
    img_embedding, txt_embedding = {}, {}
    for i in range(len(txt_ids)):
        img_embedding[img_ids[i]] = img_vecs[i]      # trainer.py:151 (last value wins, first-seen order)
        txt_embedding[txt_ids[i]] = txt_vecs[i]      # trainer.py:152
    indexer_img = flatip.FlatIndexer(txt_vecs.shape[1], scorer=scorer)
    indexer_txt = flatip.FlatIndexer(txt_vecs.shape[1], scorer=scorer)
    indexer_img.index_data(list(img_embedding.items()))
    indexer_txt.index_data(list(txt_embedding.items()))
    res_txt = indexer_img.search_knn(np.asarray(txt_vecs), num_tops)
    rank_txt_res = {txt_ids[i]: r[0] for i, r in enumerate(res_txt)}
    res_img = indexer_txt.search_knn(np.asarray(img_vecs), num_tops)
    rank_img_res = {img_ids[i]: r[0] for i, r in enumerate(res_img)}
    recall_txt = {1: 0, 5: 0, 10: 0}
    for i, q in enumerate(txt_ids):
        for top in recall_txt:
            recall_txt[top] += img_ids[i] in rank_txt_res[q][:top]
    for top in recall_txt:
        recall_txt[top] = recall_txt[top] / len(rank_txt_res)
    recall_img = {1: 0, 5: 0, 10: 0}
    for q in np.unique(img_ids):
        for top in recall_img:
            recall_img[top] += any(t in rank_img_res[q][:top] for t in img2txt[q])
    for top in recall_img:
        recall_img[top] = recall_img[top] / len(rank_img_res)
    return recall_txt, recall_img, rank_txt_res, rank_img_res
