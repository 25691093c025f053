#!/usr/bin/env python3
"""Stands in for the `ffmpeg` binary in the inference-CLI tests: understands exactly the command line the reader issues
(`-nostdin -y -i VIDEO -start_number 0 -q 0 -vf fps=F OUT/%07d.png`).  A test "video" is a .npy array of uint8 RGB frames
[n, H, W, 3] stored at 1 frame per second (whatever its file extension); frame j of the output is source frame
floor(j / F) -- enough to exercise the reader, the frame numbering and the timestamps."""
import sys

import numpy as np
from PIL import Image

args = sys.argv[1:]
video = args[args.index("-i") + 1]
fps = float(args[args.index("-vf") + 1].split("=")[1])
pattern = args[-1]
assert args[args.index("-start_number") + 1] == "0" and "-nostdin" in args and "-y" in args
frames = np.load(video, allow_pickle=False)
n_out = int(len(frames) * fps)
for j in range(n_out):
    Image.fromarray(frames[min(int(j / fps), len(frames) - 1)]).save(pattern % j)
