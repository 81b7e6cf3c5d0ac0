#!/bin/bash
# builds the timing probes next to their sources (binaries are git-ignored; they travel to the GPU box with the snapshot)
cd "$(dirname "$0")"
for f in *.cu; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o "${f%.cu}.bin" "$f" || exit 1; done
