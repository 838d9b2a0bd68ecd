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


# ------------------------------------------------------------------------------------------
# Window hand-over with the EARLY BASE of the bulk CTAs (pcd_window.cu, bulk_role / engine_role,
# horizon 0): bulk CTA b reduces the (cold) records of its positions of window w+1 while the engine
# is still in window w; the engine publishes nzwin[w] with eng_done and, when it is 0, starts window
# w+1 on the early sums without waiting for WB(w) / an exact BASE(w+1).
def _hand_over_problem(rng, n_win, nb, n_rec, p_move):
    wins = []
    for _ in range(n_win):
        recs = rng.sample(range(n_rec), nb * 2)          # every record is touched by one position per window
        wins.append([dict(recs=recs[2 * i:2 * i + 2], moves=rng.random() < p_move) for i in range(nb)])
    return wins


def _decide(w, i, p, s, state):
    upd = ((s + state + w + i) % 4 + 1) if p["moves"] else 0
    return upd, state + (upd != 0)


def _hand_over_sequential(wins, n_rec):
    rec, state, upds = list(range(n_rec)), 0, []
    for w, win in enumerate(wins):
        for i, p in enumerate(win):
            upd, state = _decide(w, i, p, sum(rec[r] for r in p["recs"]), state)
            upds.append(upd)
            for r in p["recs"]:
                rec[r] += upd
    return rec, upds


def _bulk(b, nbulk, wins, sh):
    n_win, early_valid = len(wins), False

    def base(w):
        for i in range(b, len(wins[w]), nbulk):
            s = 0
            for r in wins[w][i]["recs"]:
                s += sh["rec"][r]
                yield
            sh["base"][w][i] = s
            yield

    for w in range(n_win):
        if not early_valid:
            while (w >= 1 and sh["wb_cnt"][w - 1] < nbulk) or sh["eng_done"] < w:
                yield
            yield from base(w)
            sh["base_cnt"][w] += 1
        if w + 1 < n_win:
            yield from base(w + 1)                           # early: assumes window w changes no record
            sh["early_cnt"][w + 1] += 1
        while sh["eng_done"] < w + 1:
            yield
        early_valid = sh["nzwin"][w] == 0
        if not early_valid:
            for i in range(b, len(wins[w]), nbulk):
                upd = sh["res"][w][i]
                if upd != 0:
                    for r in wins[w][i]["recs"]:
                        sh["rec"][r] += upd
                        yield
        sh["wb_cnt"][w] += 1
        yield


def _engine(nbulk, wins, sh, out):
    state, prev_nz = 0, 1
    for w, win in enumerate(wins):
        if w > 0 and prev_nz == 0:
            while sh["early_cnt"][w] < nbulk:
                yield
        else:
            while sh["base_cnt"][w] < nbulk or (w >= 1 and sh["wb_cnt"][w - 1] < nbulk):
                yield
        nz = 0
        for i, p in enumerate(win):
            upd, state = _decide(w, i, p, sh["base"][w][i], state)
            sh["res"][w][i] = upd
            out.append(upd)
            nz += upd != 0
            yield
        sh["nzwin"][w] = nz
        prev_nz = nz
        yield
        sh["eng_done"] = w + 1


@pytest.mark.parametrize("p_move", [0.0, 0.02, 0.2, 1.0])
@pytest.mark.parametrize("seed", range(10))
def test_early_base_hand_over_matches_the_sequential_sweep(seed, p_move):
    rng = random.Random(77 * seed + int(1000 * p_move))
    n_win, nb, n_rec, nbulk = 9, 6, 40, 3
    wins = _hand_over_problem(rng, n_win, nb, n_rec, p_move)
    want_rec, want_upds = _hand_over_sequential(wins, n_rec)
    sh = dict(rec=list(range(n_rec)), base=[[None] * nb for _ in range(n_win)], res=[[None] * nb for _ in range(n_win)],
              base_cnt=[0] * n_win, early_cnt=[0] * (n_win + 1), wb_cnt=[0] * n_win, nzwin=[None] * n_win, eng_done=0)
    out = []
    actors = [_engine(nbulk, wins, sh, out)] + [_bulk(b, nbulk, wins, sh) for b in range(nbulk)]
    live, steps = list(range(len(actors))), 0
    while live:
        i = rng.choice(live)
        try:
            next(actors[i])
        except StopIteration:
            live.remove(i)
        steps += 1
        assert steps < 2_000_000, "hand-over dead-locked"
    assert out == want_upds
    assert sh["rec"] == want_rec
