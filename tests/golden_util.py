import hashlib
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_cases():
    return json.load(open(os.path.join(GOLD, "reference_outputs.json")))["cases"]


def digest(obj) -> str:
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def full_edges(edges, n, m):
    """alignment_edges(only_local=false) (alignment.cc:120-167): raw trace edges (-1 = gap) -> edges incl. locality gaps (-3)."""
    out, la, lb = [], 1, 1
    for i, j in edges:
        if i > 0:
            while la < i:
                out.append([la, -3]); la += 1
        if j > 0:
            while lb < j:
                out.append([-3, lb]); lb += 1
        if i > 0:
            la += 1
        if j > 0:
            lb += 1
        out.append([i, j])
    while la <= n:
        out.append([la, -3]); la += 1
    while lb <= m:
        out.append([-3, lb]); lb += 1
    return out
