"""eval_mode decorator with the reference's semantics (modules/utils.py:7-15)."""


def eval_mode(fn):
    def inner(self, *args, **kwargs):
        was_training = self.training
        self.eval()
        out = fn(self, *args, **kwargs)
        self.train(was_training)
        return out

    return inner
