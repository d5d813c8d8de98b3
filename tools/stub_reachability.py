"""Which Cuda_Stream methods does the adaptor NOT provide, and can any sampled toolkit reach them?  (build container only: reads
/root/reference). For every NTS_B200_UNSUPPORTED stub of sample-based-gnn_b200/host/cuda/ntsCUDA.hpp: the live (non-comment) call
sites in the reference's core/ and toolkits/, the method or op class that contains each call site, and the live uses of that
method / class in toolkits/*SAMPLE*. Prints a markdown table (pasted into INTEGRATION.md).

    python tools/stub_reachability.py
"""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("NTS_REFERENCE", "/root/reference")
ADAPTOR = os.path.join(ROOT, "sample-based-gnn_b200", "host", "cuda", "ntsCUDA.hpp")


def live_lines(path):
    """(line number, text) of the lines that are not inside // or /* */ comments"""
    out, in_block = [], False
    for i, l in enumerate(open(path, errors="replace"), 1):
        t = l
        if in_block:
            if "*/" in t:
                t = t.split("*/", 1)[1]
                in_block = False
            else:
                continue
        t = re.sub(r"/\*.*?\*/", "", t)
        if "/*" in t:
            t = t.split("/*", 1)[0]
            in_block = True
        t = t.split("//", 1)[0]
        if t.strip():
            out.append((i, t))
    return out


def enclosing(path, line_no):
    """nearest preceding `class X` or method definition"""
    best = ("?", "?")
    cls = meth = None
    for i, t in live_lines(path):
        if i > line_no:
            break
        m = re.search(r"\bclass\s+(\w+)", t)
        if m:
            cls = m.group(1)
        m = re.match(r"\s*(?:inline\s+|static\s+|virtual\s+)?[\w:<>\*&]+\s+(\w+)\s*\([^;]*$", t)
        if m and m.group(1) not in ("if", "for", "while", "switch", "return"):
            meth = m.group(1)
    return cls, meth


def count_args(text, start):
    """number of top-level arguments of the call whose '(' is at text[start]"""
    depth, n, seen = 0, 1, False
    for ch in text[start:]:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
            if depth == 0:
                return n if seen else 0
        elif ch == "," and depth == 1:
            n += 1
        elif not ch.isspace():
            seen = True
    return n


HOST_CLASSES = {"FastSampler", "NtsContext", "NtsScheduler"}   # methods are looked up by name; op classes by class name


def main():
    text = open(ADAPTOR).read()
    # unconditional stubs only (a guard like "X with feature_size != 1" is an argument check of a provided method)
    stubs = sorted({m.group(1) for m in re.finditer(r'NTS_B200_UNSUPPORTED\("(\w+)([^"]*)"', text) if " with " not in m.group(2)})
    files = sorted(glob.glob(os.path.join(REF, "core", "*.hpp")) + glob.glob(os.path.join(REF, "toolkits", "*.hpp")))
    toolkits = sorted(glob.glob(os.path.join(REF, "toolkits", "*SAMPLE*.hpp")))
    tk_text = {f: "\n".join(t for _, t in live_lines(f)) for f in toolkits}
    print("| adaptor stub | live call sites in core/ + toolkits/ | inside | live uses from toolkits/*SAMPLE* |")
    print("|---|---|---|---|")
    reachable = 0
    for s in stubs:
        sites = []
        for f in files:
            for i, t in live_lines(f):
                if re.search(r"(->|\.)" + s + r"\s*\(", t):
                    sites.append((f, i))
        if not sites:
            print(f"| `{s}` | none | - | - |")
            continue
        for f, i in sites:
            cls, meth = enclosing(f, i)
            uses = []
            if cls in HOST_CLASSES:
                for tf, tt in tk_text.items():
                    for m in re.finditer(r"(->|\.)" + re.escape(meth) + r"\s*\(", tt):
                        uses.append(f"{os.path.basename(tf)} ({count_args(tt, m.end() - 1)} args)")
            else:
                for tf, tt in tk_text.items():
                    if re.search(r"\b" + re.escape(cls) + r"\b", tt):
                        uses.append(os.path.basename(tf))
            rel = os.path.relpath(f, REF)
            note = ""
            if uses and cls == "FastSampler" and meth == "load_share_embedding":
                # overloads: the toolkits call the 7-argument form (.., super_batch_id), core/ntsFastSampler.hpp:514 -> the PROVIDED
                # 8-argument Cuda_Stream::dev_load_share_embedding; the stubbed ones are reached from the 6- and 8-argument forms only
                stub_arity = {461: 6, 591: 8}.get(i)
                if stub_arity is not None and all(f"({stub_arity} args)" not in u for u in uses):
                    note = f" -- all {len(uses)} toolkit calls use the 7-argument overload (:514, provided); this {stub_arity}-argument overload has no caller"
                    uses = []
                elif stub_arity is None:
                    note = " -- this call site is the provided overload (:514 -> nb_row_override)"
                    uses = []
            if uses:
                reachable += 1
            shown = sorted(set(uses))
            print(f"| `{s}` | `{rel}:{i}` | `{cls}::{meth}` | {', '.join(shown[:5]) if shown else 'none'}{note} |")
    print(f"\nstubs with a live path from a sampled toolkit: {reachable}")


if __name__ == "__main__":
    main()
