def custom_root(optimality_fun, has_aux=False, solve=None):
    def wrapper(solver_fun):
        def wrapped(*args, **kwargs):
            return solver_fun(*args, **kwargs)

        wrapped.optimality_fun = optimality_fun
        return wrapped

    return wrapper
