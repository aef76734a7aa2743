"""Registers / spills per kernel from the build's ptxas log: python tools/ptxas_summary.py [log]"""
import re
import sys

log = sys.argv[1] if len(sys.argv) > 1 else "aad_b200/csrc/build/ptxas_aad_kernels.log"
txt = open(log).read()
pat = (r"Compiling entry function '(\S+)'.*?\n.*?\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, "
       r"(\d+) bytes spill loads\n.*?Used (\d+) registers")
for m in re.finditer(pat, txt):
    name = re.sub(r"_ZN\d+_GLOBAL__N__\w+?_aad_kernels_cu_\w{8}\d+", "", m.group(1))
    name = re.sub(r"v18aadk_\w+_params$", "", name)
    print(f"{name[:48]:48s} stack {m.group(2):>4s} spill {m.group(3):>4s}/{m.group(4):<4s} regs {m.group(5)}")
