"""uniter_model/model/itm.py: train_itm.py:20 imports UniterForImageTextRetrieval, the cross-encoder TEACHER of the
knowledge-distillation mode (--teacher_checkpoint, train_itm.py:85-95).  The teacher is not part of the bi-encoder retrieval
path; the name exists so the script imports, and using it fails loudly."""


class UniterForImageTextRetrieval(object):
    @classmethod
    def from_pretrained(cls, *args, **kwargs):
        raise NotImplementedError("the UNITER cross-encoder teacher (knowledge distillation) is outside the retrieval hot "
                                  "path this package implements; run without --teacher_checkpoint")

    def __init__(self, *args, **kwargs):
        self.from_pretrained()
