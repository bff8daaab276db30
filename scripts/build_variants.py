#!/usr/bin/env python
"""Tuning variants of the library next to the default one: regcm_b200/variants/<name>.so (git-ignored; they
travel to the GPU box).  usage: python scripts/build_variants.py name=-DFLAG=1,-DOTHER=2 name2=..."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from regcm_b200 import build as B  # noqa: E402

vdir = os.path.join(ROOT, "regcm_b200", "variants")
os.makedirs(vdir, exist_ok=True)
for f in os.listdir(vdir):
    if f.endswith(".so"):
        os.remove(os.path.join(vdir, f))


def one(arg):
    name, flags = arg.split("=", 1)
    return B.build_library(extra_flags=[x for x in flags.split(",") if x], out=os.path.join(vdir, name + ".so"))


with ThreadPoolExecutor(max_workers=4) as ex:
    for p in ex.map(one, sys.argv[1:]):
        print(p)
