"""debug: print the requests on which aigw_embeddings_count and the oracle disagree"""
import importlib.util, json, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O
spec = importlib.util.spec_from_file_location("t", os.path.join(ROOT, "tests", "test_embeddings_count_gpu.py")); t = importlib.util.module_from_spec(spec); spec.loader.exec_module(t)
from aigw_b200 import capi
g = capi.Context(0); bpe = g.bpe_load(t.VOCAB["byte_to_id"], t.VOCAB["merges"])
r = random.Random(41); ws = t.words(r)
bodies = [t.emb_body(r, ws) for _ in range(1500)] + t.ODD
res = g.embeddings_count(bpe, bodies)
n = 0
for b, x in zip(bodies, res):
    o = O.embeddings_translate("openai", b)
    if x["status"] != 4 and int(x["status"]) != o.status:
        n += 1
        if n <= 6:
            print("MISMATCH gpu", int(x["status"]), int(x["reason"]), "oracle", o.status, o.err[:80], "len", len(b), "head", b[:70], "tail", b[-90:])
print("mismatches", n)
