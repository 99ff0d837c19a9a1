"""Executable model of the fused encoder's walk over the pair table (csrc/trie_host.h, csrc/encode2.cu):
the same lookups, in the same order, one walker, plain Python.  The CPU tests compare it with the
oracle's trie encoder (oracle/ecgb_oracle.c, lib.rs:163-190) so that the table layout and the
walk rules are pinned before the CUDA kernel runs."""
import ctypes as C

import numpy as np

from ecgbyte import _lib


class PairTable:
    def __init__(self, seq, off, ids):
        L = _lib.lib()
        seq = np.ascontiguousarray(seq, np.uint32)
        off = np.ascontiguousarray(off, np.uint64)
        ids = np.ascontiguousarray(ids, np.uint32)
        n = len(off) - 1
        n_ent = C.c_uint32(0)
        meta = np.zeros(8, np.uint32)
        cls = np.zeros(256, np.uint8)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        _lib.check(L.ecgb_pairtab_host(p(seq), p(off), p(ids), n, None, None, 0, C.byref(n_ent), p(meta), p(cls)))
        self.ent = np.zeros(n_ent.value, np.uint32)
        self.tok = np.zeros(n_ent.value, np.uint16)
        _lib.check(L.ecgb_pairtab_host(p(seq), p(off), p(ids), n, p(self.ent), p(self.tok), n_ent.value,
                                       C.byref(n_ent), p(meta), p(cls)))
        (self.root_base, self.dead_base, self.W, self.NC, self.SM, self.SE, self.n_states, self.n_used) = [int(x) for x in meta]
        self.cls = cls

    def classes(self, text):
        """bytes -> classes, SE for bytes without a class"""
        c = self.cls[np.frombuffer(bytes(text), np.uint8)].astype(np.int64)
        c[c >= self.NC] = self.SE
        return c

    def encode(self, text):
        """greedy longest match of one record, the way a walker of encode2.cu does it"""
        text = bytes(text)
        c = np.concatenate([self.classes(text), [self.SE, self.SE, self.SE]])
        ent, tok, W, SM, SE = self.ent, self.tok, self.W, self.SM, self.SE
        n = len(text)
        out = []
        pos = 0
        while pos < n:
            if c[pos] == SE:          # a byte without a class is its own token (lib.rs:155-157)
                out.append(text[pos])
                pos += 1
                continue
            base = self.root_base
            m_slot, m_flags, m_pos = -1, 0, pos
            while True:
                code = (int(c[pos]) << W) | int(c[pos + 1])
                slot = base + code
                e = int(ent[slot])
                if ((e >> 2) & 0xFFF) == code:
                    if e & 3:
                        m_slot, m_flags, m_pos = slot, e & 3, pos
                    base = e >> 16
                    pos += 2
                    continue
                # the pair failed: does the first symbol alone reach a token?
                code1 = (int(c[pos]) << W) | SM
                slot1 = base + code1
                e1 = int(ent[slot1])
                if ((e1 >> 2) & 0xFFF) == code1:
                    out.append(int(tok[slot1]))
                    pos = pos + 1
                elif m_flags & 2:
                    out.append(int(tok[m_slot]))
                    pos = m_pos + 2
                elif m_flags & 1:
                    c1 = (int(ent[m_slot]) >> 2) & ((1 << W) - 1)
                    out.append(int(tok[m_slot - c1 + SM]))
                    pos = m_pos + 1
                else:
                    raise AssertionError("no terminal on a walk from the root")
                break
        return np.array(out, np.int64)
