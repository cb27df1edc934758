"""Exhaustive interleaving check of the hand-back protocol of the persistent solve kernel
(k_fim, planning-path_planning_b200/csrc/dymu_fim.cu: `ctrl[7]`, `stop_flag`).

A streamed solve asks the running kernel to return at the next phase boundary: the copy engine
sets a device word, CTA 0 looks at it at the end of every phase and publishes the request for the
other CTAs, which read it at the top of the next iteration.  All CTAs must leave in the SAME
iteration -- one that leaves early is missing at the next grid barrier and the rest of the grid
waits for it for ever.  CTA 0 publishes BEFORE it arrives at the barrier of its phase, and after
that barrier it may run a whole phase ahead of a CTA that is slow to leave it; so a plain flag can
be seen a phase early.  The kernel therefore publishes the iteration the request applies to.

The model below runs N CTAs over a few iterations under every interleaving and every arrival time
of the device word; it must find the deadlock of the plain flag (this is what happened on the GPU)
and none for the stamped request.  CPU only: it checks the protocol, not the CUDA code."""
from collections import deque

N_CTAS = 3
MAX_ITER = 4  # iterations after which a CTA returns anyway (the list drained)

# program counter of a CTA
TOP, END, WAIT, DONE = 0, 1, 2, 3


def successors(state, stamped):
    """state = (flag_set, stop_word, counter, ((pc, j), ...)); yields every state one atomic step away."""
    flag, stop, counter, ctas = state
    if not flag:  # the copy engine may report at any moment
        yield (True, stop, counter, ctas)
    for c, (pc, j) in enumerate(ctas):
        def with_cta(new, stop_=stop, counter_=counter):
            return (flag, stop_, counter_, ctas[:c] + (new,) + ctas[c + 1:])
        if pc == TOP:
            if j >= MAX_ITER:
                yield with_cta((DONE, j))
            elif stamped and stop != 0 and stop <= j + 1:  # "leave before iteration stop - 1"
                yield with_cta((DONE, j))
            elif not stamped and stop != 0:
                yield with_cta((DONE, j))
            else:
                yield with_cta((END, j))  # (the phase's work has no effect on the protocol)
        elif pc == END:
            # CTA 0 looks at the device word and publishes, then everybody arrives at barrier j
            new_stop = stop
            if c == 0 and flag and stop == 0:
                new_stop = j + 2 if stamped else 1
            yield with_cta((WAIT, j), new_stop, counter + 1)
        elif pc == WAIT:
            if counter >= (j + 1) * N_CTAS:
                yield with_cta((TOP, j + 1))


def explore(stamped):
    """-> (states visited, deadlocked states, iterations in which CTAs of one run left)"""
    start = (False, 0, 0, tuple((TOP, 0) for _ in range(N_CTAS)))
    seen, todo = {start}, deque([start])
    deadlocks, split_exits = [], []
    while todo:
        s = todo.popleft()
        nxt = list(successors(s, stamped))
        ctas = s[3]
        live = [x for x in nxt if x[3] != ctas]  # steps of CTAs (the flag arriving is not progress)
        if not live and any(pc != DONE for pc, _ in ctas):
            deadlocks.append(s)
        if all(pc == DONE for pc, _ in ctas) and len({j for _, j in ctas}) > 1:
            split_exits.append(s)
        for x in nxt:
            if x not in seen:
                seen.add(x)
                todo.append(x)
    return len(seen), deadlocks, split_exits


def test_plain_flag_can_deadlock_the_grid():
    n, deadlocks, _ = explore(stamped=False)
    assert n > 100
    assert deadlocks, "the model no longer reproduces the defect it was written for"
    # the shape of the failure: somebody has left, somebody waits at a barrier that cannot fill
    flag, stop, counter, ctas = deadlocks[0]
    assert any(pc == DONE for pc, _ in ctas) and any(pc == WAIT for pc, _ in ctas)


def test_stamped_request_never_deadlocks_and_all_ctas_leave_together():
    n, deadlocks, split_exits = explore(stamped=True)
    assert n > 100
    assert not deadlocks
    assert not split_exits
