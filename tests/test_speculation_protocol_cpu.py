"""Executable model of the zero-update speculation protocol of the pcd window engine
(sparsepoly_b200/csrc/pcd_window.cu, "ZERO-UPDATE SPECULATION").

Not a test of the CUDA code (the GPU parity tests are): it checks the PROTOCOL -- that the worker /
chain-warp hand-shake commits, under any interleaving of the actors' shared-memory operations, exactly
what the sequential coordinate sweep computes.  The actors are generators advanced one shared-memory
operation at a time by a seeded random scheduler:

  worker(t)   snapshot (C, nz_issued) of the chain progress -> wait nz_done >= nz_issued -> per record:
              wait for the write-back flag of the last KNOWN MOVER (coordinate that starts nonzero)
              touching its slot -> read the records one by one (stale / torn reads are possible) ->
              send (sum, C) -> wait for the result; REDO: evaluate again with the chain parked on t ->
              if the update is nonzero write the records back one by one, nz_done += 1 -> raise the flag
  chain       takes the run of ready cells starting at its position, rejects the first cell whose
              snapshot predates a SURPRISE (a coordinate that started at zero and moved), evaluates the
              lanes against the current regularizer state and commits up to and including the first
              position that moves

The per-position arithmetic is an arbitrary deterministic function of (records read, state): any stale
read that the protocol lets through changes the committed value and fails the comparison.
"""
import random

import pytest

EXACT = 1 << 30


def make_window(rng, nb, n_slots, p_mover, p_surprise):
    pos = []
    for t in range(nb):
        slots = rng.sample(range(n_slots), rng.randint(1, min(6, n_slots)))
        pos.append(dict(slots=slots, coef=[rng.randint(1, 5) for _ in slots],
                        mover=rng.random() < p_mover, surprise_mod=max(1, int(1 / max(p_surprise, 1e-9)))))
    return pos


def step_value(t, p, reads, state):
    """(update, new state) of position t from the record values it read and the chain state."""
    val = sum(c * r for c, r in zip(p["coef"], reads)) + 7 * state + t
    if p["mover"]:
        upd = (val % 5) - 2                       # known movers move (sometimes by exactly 0)
    else:
        upd = (val % 3 + 1) if val % p["surprise_mod"] == 0 else 0     # surprises are rare
    return upd, state + (1 if upd != 0 else 0)


def sequential(pos, n_slots):
    rec, state, upds = [1] * n_slots, 0, []
    for t, p in enumerate(pos):
        upd, state = step_value(t, p, [rec[s] for s in p["slots"]], state)
        upds.append(upd)
        for s, c in zip(p["slots"], p["coef"]):
            rec[s] += upd * c
    return rec, state, upds


class Shared:
    def __init__(self, pos, n_slots):
        nb = len(pos)
        self.rec = [1] * n_slots
        self.prog = (0, 0)                         # (decided positions C, record-changing updates issued)
        self.nz_done = 0
        self.cell = [None] * nb                    # worker -> chain: (reads, csnap)
        self.result = [None] * nb                  # chain -> worker: upd or "REDO"
        self.wbflag = [False] * nb
        # last known mover touching each slot before position t (the engine's per-slot masks)
        self.slot_movers = [[t for t, p in enumerate(pos) if p["mover"] and s in p["slots"]] for s in range(n_slots)]


def worker(t, p, sh):
    attempt = 0
    while True:
        c_snap, nzi = sh.prog
        yield
        while sh.nz_done < nzi:
            yield
        if attempt == 0:
            for s in p["slots"]:
                movers = [m for m in sh.slot_movers[s] if m < t]
                if movers:
                    while not sh.wbflag[movers[-1]]:
                        yield
        reads = []
        for s in p["slots"]:
            reads.append(sh.rec[s])
            yield
        sh.cell[t] = (reads, c_snap if attempt == 0 else EXACT)
        yield
        while sh.result[t] is None or (attempt > 0 and sh.result[t] == "REDO"):
            yield
        if sh.result[t] == "REDO" and attempt == 0:
            attempt = 1
            continue
        upd = sh.result[t]
        break
    if upd != 0:
        for s, c in zip(p["slots"], p["coef"]):
            sh.rec[s] += upd * c
            yield
        sh.nz_done += 1
        yield
    sh.wbflag[t] = True


def chain(pos, sh, out, width=8):
    nb, tl, state, last_sur, issued, redo = len(pos), 0, 0, -1, 0, -1
    while tl < nb:
        yield
        n = 0
        while n < width and tl + n < nb and sh.cell[tl + n] is not None and not (
                tl + n == redo and sh.cell[tl + n][1] != EXACT):
            n += 1
        if n == 0:
            continue
        rej = [i for i in range(n) if last_sur >= sh.cell[tl + i][1]]
        if rej:
            if rej[0] == 0:
                sh.result[tl] = "REDO"
                redo = tl
                continue
            n = rej[0]
        for i in range(n):
            t = tl + i
            upd, new_state = step_value(t, pos[t], sh.cell[t][0], state)
            sh.result[t] = upd
            out[t] = upd
            if upd != 0:                           # first position that moves: commit it and cut the run
                state = new_state
                issued += 1
                if not pos[t]["mover"]:
                    last_sur = t
                n = i + 1
                break
        tl += n
        sh.prog = (tl, issued)
    out["state"] = state


@pytest.mark.parametrize("p_mover,p_surprise", [(0.0, 0.0), (0.0, 0.1), (0.3, 0.05), (0.55, 0.1), (1.0, 0.0)])
@pytest.mark.parametrize("seed", range(12))
def test_protocol_commits_the_sequential_sweep(seed, p_mover, p_surprise):
    rng = random.Random(1000 * seed + int(100 * p_mover) + int(1000 * p_surprise))
    nb, n_slots, n_workers = 40, 14, 5
    pos = make_window(rng, nb, n_slots, p_mover, p_surprise)
    want_rec, want_state, want_upds = sequential(pos, n_slots)
    sh, out = Shared(pos, n_slots), {}
    # worker w owns positions w, w + n_workers, ... and handles them one after the other (as a warp does)

    def worker_warp(w):
        for t in range(w, nb, n_workers):
            yield from worker(t, pos[t], sh)

    actors = [chain(pos, sh, out)] + [worker_warp(w) for w in range(n_workers)]
    live = list(range(len(actors)))
    steps = 0
    while live:
        i = rng.choice(live)
        try:
            next(actors[i])
        except StopIteration:
            live.remove(i)
        steps += 1
        assert steps < 2_000_000, "protocol dead-locked"
    assert [out[t] for t in range(nb)] == want_upds
    assert out["state"] == want_state
    assert sh.rec == want_rec
    assert all(sh.wbflag)
