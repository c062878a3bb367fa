def check_grads(*a, **k):
    raise NotImplementedError("refshim: check_grads is not provided")
