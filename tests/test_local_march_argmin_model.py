"""Model check of the two-warp minimum search of k_local_march
(planning-path_planning_b200/csrc/dymu_local.cu).

The reference takes the next node of the local wave with a strict '<' scan over a vector
(minCostLocalNode, src/DyMu_LocalPathRepairing.cpp:752-805): smallest key, earliest position on
ties.  The kernel does not scan on the critical path: while warp 0 updates the neighbours of pop
k, warp 1 scans the band for the best of the REST -- and, racing with warp 0, may read the old or
the new key of a slot that pop k lowers.  Pop k+1 is then the lexicographic (key, slot) minimum of
that answer and the at most four entries pop k lowered or created.  This test replays random waves
with many ties against the plain scan, with every lowered slot read old or new at random.
CPU only: it checks the argument, not the CUDA code."""
import random

INF = float("inf")


def plain_scan(keys):
    """strict '<', earliest position: what the reference does"""
    best, pos = INF, None
    for q, k in enumerate(keys):
        if k < best:
            best, pos = k, q
    return pos


def lexmin(cands):
    cands = [(k, q) for k, q in cands if q is not None and k < INF]
    return min(cands)[1] if cands else None


def run_wave(rng, pops, key_range):
    keys = [float(rng.randrange(key_range))]  # the agent's entry in slot 0
    # state carried from pop to pop, as the kernel does
    events = [(keys[0], 0)]         # what the previous pop lowered or created (here: the agent)
    rest = (INF, None)              # warp 1's answer for the previous pop
    for _ in range(pops):
        bp = lexmin(events + [rest])
        assert bp == plain_scan(keys), (keys, events, rest)
        if bp is None:
            return
        # ---- pop: erase (tombstone), then lower up to 4 live entries and push up to 4 new ones
        before = list(keys)
        keys[bp] = INF
        events = []
        live = [q for q, k in enumerate(keys) if k < INF]
        for q in rng.sample(live, min(len(live), rng.randrange(0, 5))):
            if keys[q] > 0 and rng.random() < 0.8:
                keys[q] = float(rng.randrange(int(keys[q])))  # strictly lower, ties with others likely
                events.append((keys[q], q))
        for _ in range(rng.randrange(0, 5 - len(events))):
            keys.append(float(rng.randrange(key_range)))
            events.append((keys[-1], len(keys) - 1))
        # ---- warp 1: best of the rest over the slots that existed when the pop was announced,
        # without the popped one; a slot being lowered is read old or new
        seen = []
        for q in range(len(before)):
            if q == bp:
                seen.append(INF)
            elif keys[q] != before[q]:
                seen.append(rng.choice((before[q], keys[q])))
            else:
                seen.append(before[q])
        q_rest = plain_scan(seen)
        rest = (seen[q_rest], q_rest) if q_rest is not None else (INF, None)


def test_combined_minimum_equals_the_plain_scan():
    rng = random.Random(20261019)
    for trial in range(300):
        run_wave(rng, pops=rng.randrange(1, 120), key_range=rng.choice((3, 8, 50, 1000)))
