"""Print registers / spills / shared memory per kernel from the ptxas logs of the last build."""
import re, subprocess, sys, os
B = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tehmm_b200", "build")
units = sys.argv[1:] or ["tile", "viterbi", "emission", "stats", "forward", "backward", "fused"]
for f in units:
    p = os.path.join(B, f + ".ptxas.log")
    if not os.path.exists(p):
        continue
    txt = open(p).read()
    for m in re.finditer(r"Compiling entry function '([^']+)'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s+ptxas info\s+: Used (\d+) registers([^\n]*)", txt, re.S):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        smem = re.search(r"(\d+) bytes smem", m.group(6))
        print("%-9s %-62s regs %3s spill %s/%s smem %s" % (f, name[:62], m.group(5), m.group(3), m.group(4), smem.group(1) if smem else "0"))
