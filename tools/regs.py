#!/usr/bin/env python
"""Summarise `ptxas -v` output in csrc/build.log: registers, spills, shared memory per kernel.
usage: python tools/regs.py [pattern]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "deep-statistical-solver-for-distribution-system-state-estimation_b200", "csrc", "build.log")


def main():
    pat = sys.argv[1] if len(sys.argv) > 1 else ""
    text = open(LOG).read()
    try:
        text = subprocess.run(["c++filt"], input=text, capture_output=True, text=True).stdout
    except OSError:
        pass
    name = None
    spill = ""
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '(.*)' for", line)
        if m:
            name = re.sub(r"\(anonymous namespace\)::", "", m.group(1))
            name = re.sub(r"\((EaArgs|Tc2Args|GwArgs|WlsArgs).*\)", "", name)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            spill = f"stack {m.group(1)} spill st/ld {m.group(2)}/{m.group(3)}"
            continue
        m = re.search(r"Used (\d+) registers", line)
        if m and name and pat in name:
            print(f"{int(m.group(1)):4d} regs  {spill:32s} {name}")
            name = None


if __name__ == "__main__":
    main()
