#!/bin/bash
# K16 wavefront kernel: resident blocks per SM beyond 8 (56 / 48 / 40 registers, with spills) -- the kernel is latency-bound at 45 % occupancy
mkdir -p gpurun_out
TAG=r02Q tools/r02_variants.sh wocc9 wocc10 wocc12
