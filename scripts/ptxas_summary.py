"""Summarise an `nvcc -Xptxas -v` log: registers / spills / smem per kernel."""
import re
import subprocess
import sys


def main(path, pattern=""):
    txt = open(path).read()
    names = re.findall(r"Compiling entry function '([^']+)'", txt)
    dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    blocks = txt.split("Compiling entry function '")[1:]
    for name, blk in zip(dem, blocks):
        short = name.replace("fftwpp_gpu::(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        if pattern and not re.search(pattern, short):
            continue
        spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        regs = re.search(r"Used (\d+) registers", blk)
        print(f"{short:48s} regs {regs.group(1):>3s}  stack {spill.group(1):>4s}  spill st/ld {spill.group(2)}/{spill.group(3)}")


if __name__ == "__main__":
    main(*sys.argv[1:])
