"""Test backend for ganon_b200.build.run_build: the device calls (K2, filter creation, insertion) answered by the oracle,
so that the orchestration, the parameter choice and the file writer can be checked on the CPU; the GPU backend must then
produce the same file byte for byte."""
import numpy as np

from ganon_b200 import synth
from oracle import oracle as O


class OracleBackend:
    def __init__(self):
        self.ibf = None

    def minimisers(self, seqs, k, w):
        parts = [O.minimiser_hash(s, k, w) for s in seqs]
        return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint64)

    def create(self, n_bins, bin_size_bits, hash_functions, k, w):
        self.ibf = O.OracleIBF(n_bins, bin_size_bits, hash_functions)

    def emplace(self, hashes, bins):
        synth.emplace_numpy(self.ibf.data, self.ibf.bin_words, self.ibf.bin_size, self.ibf.hash_funs, np.asarray(hashes, dtype=np.uint64), np.asarray(bins))

    def words(self):
        return self.ibf.data

    def close(self):
        pass
