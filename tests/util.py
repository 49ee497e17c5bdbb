import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat.json")
_kat = None


def kat():
    global _kat
    if _kat is None:
        with open(GOLDEN) as f:
            _kat = json.load(f)
    return _kat


def h(x: str) -> int:
    return int(x, 16)
