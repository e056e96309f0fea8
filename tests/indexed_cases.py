"""Shared generator for the IndexedMatch tests: descriptors of two keyframes with true correspondences and vocabulary-like
candidate lists (features falling into the same 'word')."""
import numpy as np


def make_case(nA=600, nB=700, seed=0, words=64, dup=True, shuffle=True):
    rng = np.random.default_rng(seed)
    descB = rng.integers(0, 256, (nB, 32), dtype=np.uint8)
    src = rng.integers(0, nB, nA)
    descA = descB[src].copy()
    nflip = rng.integers(0, 40, nA)
    for i in range(nA):
        pos = rng.integers(0, 256, nflip[i])
        for b in pos:
            descA[i, b // 8] ^= np.uint8(1 << (b % 8))
    fresh = rng.random(nA) < 0.2                        # A features with no counterpart
    descA[fresh] = rng.integers(0, 256, (int(fresh.sum()), 32), dtype=np.uint8)
    # vocabulary word of a descriptor: a hash of a few robust bytes; correspondences mostly share the word
    wordB = (descB[:, 0].astype(np.int64) * 7 + descB[:, 5]) % words
    wordA = wordB[src].copy()
    stray = rng.random(nA) < 0.15
    wordA[stray | fresh] = rng.integers(0, words, int((stray | fresh).sum()))
    bucketB = [np.flatnonzero(wordB == w) for w in range(words)]
    bucketA = [np.flatnonzero(wordA == w) for w in range(words)]
    def lists(word, bucket):
        out = []
        for w in word:
            l = bucket[w].copy()
            if shuffle:
                rng.shuffle(l)
            if dup and len(l) and rng.random() < 0.1:
                l = np.concatenate([l, l[:1]])            # a duplicated entry (the reference would count it twice)
            out.append(l.astype(np.int32))
        return out
    a2b = lists(wordA, bucketB)
    b2a = lists(wordB, bucketA)
    maskA = (rng.random(nA) < 0.9).astype(np.uint8)
    maskB = (rng.random(nB) < 0.9).astype(np.uint8)
    return descA, descB, a2b, b2a, maskA, maskB


def csr(lists):
    off = np.zeros(len(lists) + 1, np.int32)
    off[1:] = np.cumsum([len(l) for l in lists])
    cand = np.concatenate(lists).astype(np.int32) if off[-1] else np.zeros(0, np.int32)
    return off, cand
