"""BASELINE configs[0] at its stated size, end to end on the GPU, against the REFERENCE's own run of it (north_star:
"Recall@1/5/10 identical to the reference").

tests/golden/configs0_planted.npz was minted by oracle/make_golden.py --configs0: the unmodified reference towers
(BertEncoder / UniterEncoder, 12 layers, seeded weights) and the unmodified eval_model_on_dataloader (dvl/trainer.py:113-190)
on CPU fp32 over a database directory of 1 000 images (<= 36 regions x 2048-d) and 5 000 captions (<= 32 tokens), batch 80,
both retrieval directions.  Which image a synthetic caption "describes" is a free label; the fixture assigns the labels so
that every labelled pair sits at least `delta` (fp16: 0.3, bf16: 1.0 score units - more than twice the largest ranking-
relevant score error measured for the respective CUDA towers) away from the 1|2, 5|6, 10|11 rank boundaries in BOTH
directions (oracle/planted.py).  Under that margin a correct 16-bit pipeline must give EXACTLY the reference's
Recall@1/5/10; the assertions below are equalities, not tolerances.

What runs here: tests/itm_flow_tool.eval_flow = eval_itm.py's call sequence through the namesake packages - config JSON,
checkpoint file, database directories, load_dataset, build_dataloader (DataLoader workers + PrefetchLoader), BiEncoder on
the tcgen05 kernels, DenseFlatIndexer on the fused score + top-k kernel.
"""
import json
import os

import numpy as np
import pytest

import itm_flow_tool as flow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planted(golden_dir):
    z = np.load(os.path.join(golden_dir, "configs0_planted.npz"))
    return z, json.loads(str(z["db"])), json.loads(str(z["model"]))


@pytest.mark.parametrize("tag", ["fp16", "bf16"])
def test_configs0_recall_identical_to_reference(cuda_lib, planted, tmp_path, tag):
    z, db, mk = planted
    owner = z[f"{tag}_owner"].astype(int).tolist()
    ws = flow.make_workspace(str(tmp_path), n_img=db["n_img"], caps_per_img=db["caps_per_img"], layers=mk["layers"],
                             seed_db=db["seed"], seed_txt=mk["seed_txt"], seed_img=mk["seed_img"], txt2img=owner,
                             batch_size=mk["batch_size"], n_workers=4, seq_len=db["seq_len"], num_bb=db["num_bb"])
    out = flow.eval_flow(ws["config"], ws["checkpoint"], fp16=(tag == "fp16"))
    n_cap = db["n_img"] * db["caps_per_img"]
    assert len(out["rank_txt"]) == n_cap and out["n_indexed"] == len(set(owner))
    got_txt = [out["recall_txt"][t] for t in (1, 5, 10)]
    got_img = [out["recall_img"][t] for t in (1, 5, 10)]
    want_txt, want_img = z[f"{tag}_recall_txt"].tolist(), z[f"{tag}_recall_img"].tolist()
    # ranking agreement beyond what the labels pin (reported, and bounded loosely: unplanted ranks may swap inside the noise)
    ref_top = z[f"{tag}_rank_txt_top10"]
    got_top = np.array([[int(n[4:11]) for n in out["rank_txt"][str(j)][:10]] for j in range(n_cap)])
    top1 = float((got_top[:, 0] == ref_top[:, 0]).mean())
    overlap = float(np.mean([len(set(a) & set(b)) / 10 for a, b in zip(got_top, ref_top)]))
    cls = z[f"{tag}_class"]
    planted_top1 = float((got_top[cls == 0, 0] == ref_top[cls == 0, 0]).mean())
    msg = (f"{tag}: recall_txt {got_txt} vs reference {want_txt}; recall_img {got_img} vs {want_img}; loss {out['loss']:.5f} vs "
           f"{float(z[f'{tag}_loss']):.5f}; acc {out['acc']:.4f} vs {float(z[f'{tag}_acc']):.4f}; top-1 equal {top1:.4f} "
           f"(planted hit@1 captions: {planted_top1:.4f}); top-10 overlap {overlap:.4f}")
    print(msg)
    assert got_txt == want_txt, msg            # text -> image Recall@1/5/10: IDENTICAL
    assert got_img == want_img, msg            # image -> text Recall@1/5/10: IDENTICAL
    assert planted_top1 == 1.0, msg            # every caption planted as a margin-safe top-1 hit retrieves the same image
    assert abs(out["loss"] - float(z[f"{tag}_loss"])) <= (5e-3 if tag == "fp16" else 0.15), msg   # (bf16 score noise, sigma 0.2, moves the in-batch NLL)
    assert abs(out["acc"] - float(z[f"{tag}_acc"])) <= 0.003, msg
    assert top1 >= (0.98 if tag == "fp16" else 0.93) and overlap >= (0.97 if tag == "fp16" else 0.88), msg
