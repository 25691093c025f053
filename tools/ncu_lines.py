"""Attribute ncu per-SASS-instruction counters to CUDA source lines (dev tool).

usage: ncu_lines.py <report.ncu-rep> <lib.so> <kernel-name-substring> [top]
Needs ncu, cuobjdump and nvdisasm on PATH; the library must be built with -lineinfo.
"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_lines(lib, kernel_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        text = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_line = None, None
        for ln in text.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn:
                out.setdefault(cur_fn, {})[int(m.group(1), 16)] = (cur_line, m.group(2))
    return out


def main():
    rep, lib, sub = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    text = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    fn_maps = sass_lines(lib, sub)
    kernel = None
    hdr = None
    per_line = {}
    total = {"inst": 0, "samp": 0}
    base = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            kernel = r[1]
            hdr = None
            continue
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) - 2:
            continue
        d = dict(zip(hdr, r))
        try:
            addr = int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"])
        except ValueError:
            continue
        if base is None:
            base = addr
        off = addr - base
        inst = int(d.get("Instructions Executed") or 0)
        samp = int(d.get("# Samples") or 0)
        # pick the function map whose mangled name matches the kernel template args best
        cand = [k for k in fn_maps if sub in k]
        key = cand[0] if cand else None
        line = fn_maps.get(key, {}).get(off, (None, ""))[0] if key else None
        e = per_line.setdefault(line, [0, 0])
        e[0] += inst
        e[1] += samp
        total["inst"] += inst
        total["samp"] += samp
    print(f"kernel: {kernel}\ntotal warp-instructions {total['inst']}  samples {total['samp']}")
    src_cache = {}
    for line, (inst, samp) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        txt = ""
        if line and line[0]:
            path = glob.glob(os.path.join(os.path.dirname(lib), "**", line[0]), recursive=True) or \
                glob.glob(os.path.join(os.path.dirname(lib), line[0]))
            if path:
                src_cache.setdefault(path[0], open(path[0]).read().splitlines())
                if line[1] - 1 < len(src_cache[path[0]]):
                    txt = src_cache[path[0]][line[1] - 1].strip()
        print(f"{inst / max(total['inst'], 1) * 100:5.1f}% inst  {samp / max(total['samp'], 1) * 100:5.1f}% stall-samples  "
              f"{line}  {txt[:100]}")


if __name__ == "__main__":
    main()
