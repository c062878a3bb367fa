def while_loop(cond_fun, body_fun, init_val, maxiter, unroll=False, jit=True):
    """Bounded while loop: run body while cond holds, at most maxiter times."""
    val = init_val
    for _ in range(maxiter):
        if not bool(cond_fun(val)):
            break
        val = body_fun(val)
    return val
