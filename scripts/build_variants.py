"""Build variants of libnsf_b200.so for kernel experiments: python scripts/build_variants.py name:DEF=V,DEF2=V2 ...
-> scripts/_bin/libnsf_<name>.so (select with NSF_B200_LIB=...)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nsfnet_b200 import build as b

os.makedirs(os.path.join(b.ROOT, "scripts", "_bin"), exist_ok=True)


def one(spec):
    name, _, defs = spec.partition(":")
    out = os.path.join(b.ROOT, "scripts", "_bin", f"libnsf_{name}.so")
    return b.build_cuda(force=True, out=out, defines=[d for d in defs.split(",") if d])


with ThreadPoolExecutor(4) as ex:
    for p in ex.map(one, sys.argv[1:]):
        print("built", p)
