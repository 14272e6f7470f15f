// Empty stand-in for cuda-samples' helper_cuda.h, which the reference includes
// (cuda/bench_utils.cuh:4, cuda/main.cu:51) but never uses on the conv path.
// Needed only to compile the reference's own sources into oracle/_ref/ (see oracle/Makefile).
#pragma once
