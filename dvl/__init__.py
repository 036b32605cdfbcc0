"""`dvl` - namesake of the reference's top-level package (intersun/LightningDOT, dvl/), so that eval_itm.py,
train_itm.py and rerank.py run UNCHANGED against the B200 implementation:

    dvl.models.bi_encoder      -> lightningdot_b200.bi_encoder      dvl.trainer  -> lightningdot_b200.trainer
    dvl.indexer.faiss_indexers -> lightningdot_b200.indexer         dvl.utils    -> lightningdot_b200.utils
    dvl.data.itm               -> lightningdot_b200.data            dvl.options  -> lightningdot_b200.options
    dvl.hn                     -> lightningdot_b200.hn              dvl.const    (IMG_DIM ...)

Each sub-module is an alias (sys.modules entry) of its mirror.  Importing `dvl` also registers the module path
`transformers.tokenization_bert` that the scripts import BertTokenizer from (transformers 2.3.0 layout; see
lightningdot_b200/compat.py).  Put the repository root on sys.path AHEAD of the reference checkout, e.g.

    python -m lightningdot_b200.run_script /path/to/LightningDOT/eval_itm.py config.json checkpoint.pt
"""
from lightningdot_b200 import compat as _compat

_compat.install()
