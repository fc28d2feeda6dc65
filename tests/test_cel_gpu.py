"""C2 (CEL) parity: cost expressions compiled to device bytecode vs the oracle's tree interpreter and the reference's vectors."""
import numpy as np
import pytest

import _oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gw():
    from aigw_b200 import capi
    g = capi.Context(0)
    yield g
    g.close()


def usages(rows):
    from aigw_b200 import capi
    u = np.zeros(len(rows), dtype=capi.SseResult)
    for i, r in enumerate(rows):
        u[i]["input"], u[i]["cached"], u[i]["cache_creation"], u[i]["output"], u[i]["total"], u[i]["reasoning"] = r
        u[i]["mask"] = 63
    return u


def run(gw, exprs, rows, models=None, model="m", backend="b", route="r"):
    progs, rcs = [], []
    for e in exprs:
        rc, h = gw.cost_compile(e)
        rcs.append(rc); progs.append(h)
    live = [p for p in progs if p is not None]
    costs, errs = gw.usage_costs_cel_host(usages(rows), live, models=models, model=model, backend=backend, route=route) if live else (None, None)
    for p in live: gw.cost_free(p)
    return rcs, costs, errs


def test_reference_vectors(gw):
    """internal/llmcostcel/cel_test.go:15-90 and examples/token_ratelimit/token_ratelimit.yaml:49-65"""
    e1 = "model == 'cool_model' ?  (input_tokens - cached_input_tokens - cache_creation_input_tokens) * output_tokens  : total_tokens"
    rcs, costs, errs = run(gw, [e1, "int(input_tokens) - int(output_tokens)", "input_tokens - output_tokens", "output_tokens + reasoning_tokens",
                                "model == 'cool_model' ?  input_tokens * output_tokens : total_tokens", "input_tokens == uint(3) ? 100000000 : 0", "1 + 1", "uint(1) + uint(1)"],
                           [(200, 100, 1, 2, 3, 0), (100, 0, 0, 2000, 3, 0), (0, 0, 0, 100, 0, 50), (100, 0, 0, 2, 3, 0), (3, 0, 0, 0, 0, 0)],
                           models=[b"cool_model", b"cool_model", b"cool_model", b"cool_model", b"not_cool_model"], backend="cool_backend", route="cool_route")
    assert rcs == [0] * 8
    assert (int(errs[0, 0]), int(costs[0, 0])) == (0, 198)
    assert int(errs[1, 1]) == 5 and int(errs[1, 2]) == 2            # "result is negative (-1900)", "unsigned integer overflow"
    assert (int(errs[2, 3]), int(costs[2, 3])) == (0, 150)
    assert (int(errs[3, 4]), int(costs[3, 4])) == (0, 200)
    assert int(costs[4, 0]) == 0 and int(costs[4, 4]) == 0           # not_cool_model → total_tokens (0 in that row)
    assert int(costs[4, 5]) == 100000000 and int(costs[0, 5]) == 0
    assert int(costs[0, 6]) == 2 and int(costs[0, 7]) == 2
    assert gw.cost_compile("1 +")[0] == -2 and gw.cost_compile("uint(1)-uint(1200)")[0] == -3
    assert gw.cost_compile("model.startsWith('gpt') ? 1 : 2")[0] == -2 and gw.cost_compile("input_tokens * 2")[0] == -2 and gw.cost_compile("model")[0] == -3


def gen(rng, ty, depth):
    """random well-typed expression of type ty in {'u','i','b'}"""
    toks = ["input_tokens", "cached_input_tokens", "cache_creation_input_tokens", "output_tokens", "total_tokens", "reasoning_tokens"]
    if depth <= 0 or rng.integers(0, 5) == 0:
        if ty == "u": return [toks[int(rng.integers(0, 6))], "%du" % int(rng.integers(0, 5)), "%du" % int(rng.integers(0, 2**40)), "18446744073709551615u", "uint(%d)" % int(rng.integers(0, 100))][int(rng.integers(0, 5))]
        if ty == "i": return ["%d" % int(rng.integers(0, 7)), "%d" % int(rng.integers(0, 2**62)), "9223372036854775807", "int(%s)" % toks[int(rng.integers(0, 6))], "0x%x" % int(rng.integers(0, 4096))][int(rng.integers(0, 5))]
        strs = ["model", "backend", "route_name", "'gpt-4o'", "'aws'", "'r1'", '"cool_model"']
        return ["true", "false", "(%s %s %s)" % (strs[int(rng.integers(0, 7))], ["==", "!="][int(rng.integers(0, 2))], strs[int(rng.integers(0, 7))])][int(rng.integers(0, 3))]
    k = int(rng.integers(0, 6))
    if ty in "ui":
        if k <= 2: return "(%s %s %s)" % (gen(rng, ty, depth - 1), "+-*/%"[int(rng.integers(0, 5))], gen(rng, ty, depth - 1))
        if k == 3: return "(%s ? %s : %s)" % (gen(rng, "b", depth - 1), gen(rng, ty, depth - 1), gen(rng, ty, depth - 1))
        if k == 4: return ("uint(%s)" if ty == "u" else "int(%s)") % gen(rng, "i" if ty == "u" else "u", depth - 1)
        return "-(%s)" % gen(rng, "i", depth - 1) if ty == "i" else gen(rng, "u", depth - 1)
    if k <= 1: t = "ui"[int(rng.integers(0, 2))]; return "(%s %s %s)" % (gen(rng, t, depth - 1), ["==", "!=", "<", "<=", ">", ">="][int(rng.integers(0, 6))], gen(rng, t, depth - 1))
    if k == 2: return "(%s && %s)" % (gen(rng, "b", depth - 1), gen(rng, "b", depth - 1))
    if k == 3: return "(%s || %s)" % (gen(rng, "b", depth - 1), gen(rng, "b", depth - 1))
    if k == 4: return "!(%s)" % gen(rng, "b", depth - 1)
    return "(%s == %s)" % (gen(rng, "b", depth - 1), gen(rng, "b", depth - 1))


def test_random_expressions_match_oracle(gw):
    rng = np.random.default_rng(11)
    rows = [(0, 0, 0, 0, 0, 0), (1, 0, 0, 1, 2, 0), (200, 100, 1, 2, 3, 0), (4294967295,) * 6, (100, 0, 0, 2000, 3, 0), (7, 3, 2, 5, 12, 1)] + \
           [tuple(int(x) for x in rng.integers(0, 2**32, 6)) for _ in range(10)] + [tuple(int(x) for x in rng.integers(0, 50, 6)) for _ in range(16)]
    models = [[b"gpt-4o", b"cool_model", b"", b"aws"][i % 4] for i in range(len(rows))]
    n_ok = n_rej = 0
    for batch in range(40):
        exprs = [gen(rng, "ui"[int(rng.integers(0, 2))], int(rng.integers(1, 5))) for _ in range(8)]
        oc = [O.Cel(e) for e in exprs]
        rcs, costs, errs = run(gw, exprs, rows, models=models, backend="aws", route="r1")
        col = 0
        for e, c, rc in zip(exprs, oc, rcs):
            if rc == -2:
                assert c.rc == 1 or len(e) > 400, (e, c.rc)   # a type error in both, or the compiler's size limits
                continue
            assert (rc, c.rc) in ((0, 0), (-3, 2)), (e, rc, c.rc)
            if rc != 0:
                n_rej += 1; continue
            for i, r in enumerate(rows):
                ee, ev = c.eval(models[i].decode(), "aws", "r1", *r)
                assert (int(errs[i, col]), int(costs[i, col])) == (ee, ev if ee == 0 else 0), (e, r, models[i], int(errs[i, col]), int(costs[i, col]), ee, ev)
            col += 1; n_ok += 1
    assert n_ok > 100 and n_rej > 5


def test_uniform_model_and_many_records(gw):
    rng = np.random.default_rng(3)
    rows = [tuple(int(x) for x in rng.integers(0, 100000, 6)) for _ in range(50000)]
    exprs = ["model == 'gpt-4o' && backend != 'azure' ? input_tokens * 3u + output_tokens * 15u : total_tokens", "(input_tokens - cached_input_tokens) + cached_input_tokens / 4u + output_tokens * 4u"]
    rcs, costs, errs = run(gw, exprs, rows, model="gpt-4o", backend="aws", route="r")
    assert rcs == [0, 0]
    a = np.array(rows, dtype=np.uint64)
    assert (costs[:, 0] == a[:, 0] * 3 + a[:, 3] * 15).all() and (errs[:, 0] == 0).all()
    ok = a[:, 0] >= a[:, 1]
    assert (errs[:, 1] == np.where(ok, 0, 2)).all()
    assert (costs[ok, 1] == (a[ok, 0] - a[ok, 1]) + a[ok, 1] // 4 + a[ok, 3] * 4).all()
