"""Executable model (NumPy / plain Python, CPU) of the sharded training protocol that
ecg-byte_b200/csrc/train.cu + ecgbyte/dist_train.py implement on GPUs: boundary records,
halo construction (make_halo), shard-local merge with the histogram patches, and the
replicated-histogram argmax.  Test infrastructure: it lets the world_size-2 `gloo` tests
check, without a GPU, that the protocol reproduces single-string training exactly
(reference semantics: lib.rs:10-48, 85-117)."""

SENT = 0xFFFF


def boundary(tok, a, same):
    """What a rank publishes for the pair whose left token is `a`."""
    n = len(tok)
    rec = {"n": n, "first": [tok[i] if i < n else SENT for i in range(3)],
           "last": [tok[n - 2] if n >= 2 else SENT, tok[n - 1] if n >= 1 else SENT],
           "trail_par": 0, "all_a": 0}
    if same:
        run = 0
        while run < n and tok[n - 1 - run] == a:
            run += 1
        rec["trail_par"] = run & 1
        rec["all_a"] = int(run == n)
    return rec


def make_halo(all_bd, rank, a, b):
    """Left 2 / right 3 context tokens of the global stream and the parity of the run of
    `a` that ends just before this shard (mirrors ecgb::make_halo)."""
    world = len(all_bd)
    got = []
    for r in range(rank - 1, -1, -1):
        if len(got) >= 2:
            break
        n = all_bd[r]["n"]
        if n >= 1:
            got.append(all_bd[r]["last"][1])
        if n >= 2 and len(got) < 2:
            got.append(all_bd[r]["last"][0])
    L = [SENT, SENT]
    if len(got) >= 1:
        L[1] = got[0]
    if len(got) >= 2:
        L[0] = got[1]
    R = []
    for r in range(rank + 1, world):
        for i in range(min(3, all_bd[r]["n"])):
            if len(R) < 3:
                R.append(all_bd[r]["first"][i])
    nr = len(R)
    R += [SENT] * (3 - nr)
    par = 0
    if a == b:
        for r in range(rank - 1, -1, -1):
            n = all_bd[r]["n"]
            if n == 0:
                continue
            if all_bd[r]["all_a"]:
                par ^= n & 1
                continue
            par ^= all_bd[r]["trail_par"]
            break
    return {"nl": len(got), "nr": nr, "L": L, "R": R, "par_in": par}


def count_local(tok, right_tok):
    """get_stats of the shard, plus the window that straddles into the next shard."""
    d = {}
    n = len(tok)
    for i in range(n):
        r = tok[i + 1] if i + 1 < n else right_tok
        if r == SENT:
            continue
        k = (tok[i], r)
        d[k] = d.get(k, 0) + 1
    return d


def merge_local(tok, a, b, z, h):
    """Shard-local merge of (a, b) -> z with halo h; returns (new tokens, histogram patches)."""
    n = len(tok)
    same = a == b

    def at(p):
        if 0 <= p < n:
            return tok[p]
        if p < 0:
            return h["L"][2 + p] if -p <= h["nl"] else SENT
        q = p - n
        return h["R"][q] if q < h["nr"] else SENT

    site = [False] * n
    removed = [False] * n
    rs = -h["par_in"]  # virtual start of the run of a that is open at position 0
    for p in range(n):
        if same:
            isa = tok[p] == a
            odd = ((p - rs) & 1) != 0
            site[p] = isa and not odd and at(p + 1) == a
            removed[p] = isa and odd
            if not isa:
                rs = p + 1
        else:
            site[p] = tok[p] == a and at(p + 1) == b
            removed[p] = at(p - 1) == a and tok[p] == b
    delta = {}

    def add(k, v):
        delta[k] = delta.get(k, 0) + v

    for p in range(n):
        if not site[p]:
            continue
        tm2, tm1, tp2, tp3 = at(p - 2), at(p - 1), at(p + 2), at(p + 3)
        has_left = tm1 != SENT
        prev_site = has_left and tm2 == a and tm1 == b
        has_right = tp2 != SENT
        next_site = has_right and tp2 == a and tp3 == b
        add((a, b), -1)
        if has_left:
            add((tm1, a), -1)
            add((z if prev_site else tm1, z), +1)
        if has_right and not next_site:
            add((b, tp2), -1)
            add((z, tp2), +1)
    out = [z if site[p] else tok[p] for p in range(n) if not removed[p]]
    return out, {k: v for k, v in delta.items() if v != 0}


def argmax(hist):
    """lib.rs:92-94 with the deterministic rule: max count, then smallest (left, right)."""
    best, cnt, tied = None, 0, 0
    for k, c in hist.items():
        if c <= 0:
            continue
        if c > cnt:
            best, cnt, tied = k, c, 1
        elif c == cnt:
            tied += 1
            if k < best:
                best = k
    return best, cnt, tied


def train_rank(tok, num_merges, rank, world, all_gather):
    """One rank's loop.  all_gather(obj) -> list of every rank's obj, in rank order."""
    tok = list(tok)
    hist = {}
    bds = all_gather(boundary(tok, SENT, False))
    right = SENT
    for r in range(rank + 1, world):
        if bds[r]["n"]:
            right = bds[r]["first"][0]
            break
    lists = all_gather(count_local(tok, right))
    merges = []
    for step in range(num_merges):
        for lst in lists:  # commit: every rank applies every list -> identical histograms
            for k, v in lst.items():
                hist[k] = hist.get(k, 0) + v
        best, cnt, tied = argmax(hist)
        if best is None:
            break
        a, b = best
        bds = all_gather(boundary(tok, a, a == b))
        tok, delta = merge_local(tok, a, b, 256 + step, make_halo(bds, rank, a, b))
        lists = all_gather(delta)
        merges.append((a, b, cnt, tied))
    return tok, merges
