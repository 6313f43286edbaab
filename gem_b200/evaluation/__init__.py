"""Reconstruction evaluation of GEM (gem/evaluation/*) with the n^2 work on the GPU."""
