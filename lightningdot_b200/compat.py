"""Third-party module paths the reference's scripts import that moved or need the network since they were written.

  transformers.tokenization_bert   eval_itm.py:25, train_itm.py:21 (transformers==2.3.0 layout, DVL.yml:180).  Registered
                                   as an alias module whose BertTokenizer.from_pretrained resolves a hub id
                                   ('bert-base-cased', eval_itm.py:54) from the local cache when a vocabulary is there
                                   and otherwise returns OfflineBertTokenizer: the special-token ids of bert-base-cased
                                   ([CLS] 101, [SEP] 102, [MASK] 103 - the ids the text databases are built with,
                                   meta.json) and an encode() that fails loudly.  The tokenizer is only exercised by
                                   img_meta captions (dvl/data/itm.py:95-118) and retrieve_query; evaluation and training
                                   from tokenised databases never call it (hard part 9).
"""
import logging
import sys
import types

logger = logging.getLogger()


class OfflineBertTokenizer(object):
    cls_token_id, sep_token_id, mask_token_id, pad_token_id, unk_token_id = 101, 102, 103, 0, 100
    vocab_size = 28996

    def __init__(self, name):
        self.name_or_path = name

    def encode(self, text, add_special_tokens=True, **kw):
        raise RuntimeError(f"no vocabulary for '{self.name_or_path}' is available offline: put vocab.txt in a directory and "
                           "pass that directory as txt_model_config / tokenizer path")

    __call__ = tokenize = convert_tokens_to_ids = encode


def _real_bert_tokenizer():
    try:
        from transformers.models.bert.tokenization_bert import BertTokenizer
        return BertTokenizer
    except Exception:  # pragma: no cover
        return None


class BertTokenizer(object):
    """from_pretrained(name_or_dir) -> the installed transformers' BertTokenizer when it finds a real vocabulary,
    else OfflineBertTokenizer."""

    @classmethod
    def from_pretrained(cls, name, *args, **kwargs):
        real = _real_bert_tokenizer()
        if real is not None:
            try:
                kwargs.setdefault("local_files_only", True)
                tok = real.from_pretrained(name, *args, **kwargs)
                if getattr(tok, "vocab_size", 0) >= 1000:   # (transformers 5 builds an EMPTY vocabulary when no file is found)
                    return tok
            except Exception as e:  # no cached files
                logger.info("BertTokenizer.from_pretrained(%s) unavailable offline (%s)", name, type(e).__name__)
        return OfflineBertTokenizer(name)


def install():
    """Register `transformers.tokenization_bert` (idempotent; an existing module of that name is left alone)."""
    name = "transformers.tokenization_bert"
    if name in sys.modules:
        return
    try:
        import transformers
    except Exception:  # pragma: no cover
        transformers = None
    mod = types.ModuleType(name)
    mod.BertTokenizer = BertTokenizer
    mod.__doc__ = "alias registered by lightningdot_b200.compat (transformers 2.x module path)"
    sys.modules[name] = mod
    if transformers is not None:
        try:
            setattr(transformers, "tokenization_bert", mod)
        except Exception:
            pass
