import sys, collections
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import aigw_b200 as A, _workload as W, _oracle as O
ctx = A.Context(0)
bodies = [W.diverse_body(s) for s in range(4000)]
got = ctx.chat_translate(ctx.cfg("aws-bedrock"), bodies)
c = collections.Counter(); ex = {}
for b, g in zip(bodies, got):
    if g["status"] == A.AIGW_DECLINED:
        o = O.chat_translate("aws-bedrock", b, prefix="v1")
        key = (g["reason"], o.status)
        c[key] += 1; ex.setdefault(key, b[:260])
print(sum(1 for g in got if g["status"] == 0), "ok of", len(got))
for k, v in c.most_common(): print(k, v, ex[k])
