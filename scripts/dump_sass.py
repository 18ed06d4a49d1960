"""SASS listings + static opcode histograms of the kernels the roofline and the proof-time split are quoted on:
    python scripts/dump_sass.py            # reads luminair_b200/libluminair_b200.so, writes profiles/sass/"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
LIB = os.path.join(ROOT, "luminair_b200", "libluminair_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
# mangled-name fragments of the kernels to keep (first match of each)
WANT = ["cfft_low_fastILb1ELi12ELi2E", "cfft_low_fastILb0ELi12ELi2E", "cfft_high_vecILb1ELi8ELi12ELi0ELi2E",
        "cfft_high_vecILb0ELi8ELi12ELi0ELi2E", "merkle_layer_small_kernelILi2E", "merkle_subtree_kernel", "merkle_top_kernel",
        "19eval_partial_kernel", "16quotients_kernelILi1E", "16quotients_kernelILi2E", "constraint_quotients_kernelILi0E",
        "fold_kernel_dev_alphaILb1E", "fri_tail_kernel"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    name, buf = None, []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                funcs[name] = buf
            name, buf = m.group(1), []
        elif name:
            buf.append(line)
    if name:
        funcs[name] = buf
    os.makedirs(OUT, exist_ok=True)
    for f in os.listdir(OUT):
        if f.endswith(".sass"):
            os.remove(os.path.join(OUT, f))
    lines = ["SASS listings (`cuobjdump -sass luminair_b200/libluminair_b200.so`, sm_100a; regenerate with",
             "`python scripts/dump_sass.py`) of the kernels the roofline and the proof-time split are quoted on.  Static opcode",
             "histogram per kernel (top 14):", ""]
    tma = False
    for frag in WANT:
        hit = [n for n in funcs if frag in n]
        if not hit:
            print("missing", frag, file=sys.stderr)
            continue
        n = sorted(hit, key=len)[0]
        body = funcs[n]
        ops = collections.Counter()
        for l in body:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
            if m:
                ops[m.group(1)] += 1
        tma = tma or any(o.startswith(("UTMALDG", "UBLKCP")) for o in ops)
        short = re.sub(r"^_ZN2lb\d*", "", n)[:60]
        fname = "r2_" + re.sub(r"^\d+", "", re.sub(r"[^A-Za-z0-9_]", "", frag)) + ".sass"
        with open(os.path.join(OUT, fname), "w") as f:
            f.write("Function : " + n + "\n" + "\n".join(body) + "\n")
        total = sum(ops.values())
        lines.append(f"* `{fname}` ({short}...): {total} instructions; " + ", ".join(f"{o} {c}" for o, c in ops.most_common(14)))
    lines += ["", "`ACQBULK` is `griddepcontrol.wait`, `PREEXIT` `griddepcontrol.launch_dependents` (csrc/launch.cuh).",
              ("TMA instructions present." if tma else
               "No UTMALDG / UBLKCP (no TMA): tiles are staged through shared memory with 128-bit LDG/STS (north_star allows either).")]
    open(os.path.join(OUT, "README.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
