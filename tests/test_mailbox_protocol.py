"""Exhaustive interleaving check of the mailbox protocol of the cross-rank sum.

The peer-memory step (base_b200/csrc/vshard.cu, and the fused tail in lse.cu:pull_and_total)
has no barrier, no fence and no separate flag: a packet carries its step number, mailboxes
are double-buffered by step parity, and a per-chain counter in device memory names the step.
Its header claims that two parities are enough because "a rank can be at most one step
ahead of a peer".  A GPU test only ever sees the interleavings the hardware happens to
produce; this test walks ALL of them on a small model and checks that claim, and that the
checker has teeth: with ONE parity it must find the overwrite.

The model, per rank and chain, is the thread (or finishing warp) of one step:
    read    step = seq[chain] + 1
    push    one store per destination rank:  mail[dst][step & (P-1)][src][chain] = step
    pull    one load per awaited source, in any order, repeated until it shows `step`;
            a load that shows a LATER step is the failure (the packet was overwritten
            before it was read: on the device that chain would spin until the timeout)
    finish  seq[chain] = step
Stores and loads are single actions because the device's are: each 8-byte half of a packet is
atomic and carries the step.  Two orderings of a rank's steps are explored:
    stream  step s+1 of any chain starts after step s of EVERY chain of the rank has
            finished (launches ordered on one stream — the documented contract);
    chain   only a chain's own steps are ordered (what a fused kernel with programmatic
            dependent launch, or per-chain streams, would give).
"""
from __future__ import annotations

from collections import deque

import pytest

READ, PUSH, PULL, DONE = range(4)


def explore(world: int, chains: int, steps: int, parities: int, order: str, self_via_mailbox: bool):
    """BFS over every interleaving.  Returns (n_states, violation or None, worst lead): the lead is
    how many launches one rank's chain is ahead of the same chain on another rank."""
    idx = {}                                           # (dst, par, src, chain) -> position in the mail tuple
    for dst in range(world):
        for par in range(parities):
            for src in range(world):
                for c in range(chains):
                    idx[(dst, par, src, c)] = len(idx)

    def dests(r):                                      # remote peers first, self last (vshard.cu)
        d = [(r + k) % world for k in range(1, world)]
        return d + [r] if self_via_mailbox else d

    def sources(r):
        return frozenset(range(world)) if self_via_mailbox else frozenset(x for x in range(world) if x != r)

    # a thread is (phase, step, k): k = next destination index while pushing, awaited sources while pulling
    thread0 = (READ, 0, 0)
    start = (tuple([0] * len(idx)), tuple([0] * (world * chains)),
             tuple([thread0] * (world * chains)), tuple([1] * (world * chains)))   # mail, seq, threads, launch no.
    seen, todo, lead = {start}, deque([start]), 0
    while todo:
        mail, seq, threads, launch = todo.popleft()
        moved = False
        for c in range(chains):
            at = [launch[r * chains + c] for r in range(world)]
            lead = max(lead, max(at) - min(at))
        for r in range(world):
            for c in range(chains):
                t = r * chains + c
                phase, step, k = threads[t]
                nmail, nseq, nthreads, nlaunch = mail, seq, None, launch
                if phase == READ:
                    if launch[t] > steps:
                        continue                       # this chain has run all its steps
                    if order == "stream" and any(launch[r * chains + c2] < launch[t] for c2 in range(chains)):
                        continue                       # an earlier launch of this rank has not retired
                    nthreads = (PUSH, seq[t] + 1, 0)
                elif phase == PUSH:
                    d = dests(r)
                    if k < len(d):
                        m = list(mail)
                        m[idx[(d[k], step % parities, r, c)]] = step
                        nmail = tuple(m)
                        nthreads = (PUSH, step, k + 1)
                    else:
                        nthreads = (PULL, step, sources(r))
                elif phase == PULL:
                    if not k:
                        s = list(seq)
                        s[t] = step
                        nseq = tuple(s)
                        l = list(launch)
                        l[t] += 1
                        nlaunch = tuple(l)
                        nthreads = thread0
                    else:
                        for src in k:                  # any awaited slot may be the one that is read next
                            got = mail[idx[(r, step % parities, src, c)]]
                            if got > step:
                                return len(seen), (f"rank {r} chain {c} step {step}: slot of rank {src} already "
                                                   f"holds step {got}"), lead
                            if got == step:
                                th = list(threads)
                                th[t] = (PULL, step, k - {src})
                                nxt = (mail, seq, tuple(th), launch)
                                moved = True
                                if nxt not in seen:
                                    seen.add(nxt)
                                    todo.append(nxt)
                        continue
                th = list(threads)
                th[t] = nthreads
                nxt = (nmail, nseq, tuple(th), nlaunch)
                moved = True
                if nxt not in seen:
                    seen.add(nxt)
                    todo.append(nxt)
        if not moved and any(l <= steps for l in launch):
            return len(seen), f"deadlock with launches {launch}, seq {seq}", lead
        if not moved and any(s != steps for s in seq):
            return len(seen), f"finished with step counters {seq}, expected {steps} everywhere", lead
    return len(seen), None, lead


@pytest.mark.parametrize("order", ["stream", "chain"])
@pytest.mark.parametrize("world,chains,steps,self_via_mailbox", [
    (2, 1, 4, True),       # vshard_step_kernel: every rank also stores into its own mailbox
    (2, 2, 3, True),
    (2, 2, 3, False),      # fused tail: local shards come from partials[], only peers get packets
    (3, 1, 3, False),
    (3, 1, 3, True),
])
def test_two_parities_are_enough_under_every_interleaving(world, chains, steps, self_via_mailbox, order):
    n, bad, lead = explore(world, chains, steps, parities=2, order=order, self_via_mailbox=self_via_mailbox)
    assert bad is None, bad
    assert n > 100                                     # the walk did branch
    assert lead == 1                                   # "at most one step ahead of a peer" — and it does get ahead


@pytest.mark.parametrize("order", ["stream", "chain"])
@pytest.mark.parametrize("self_via_mailbox", [True, False])
def test_one_parity_is_caught(order, self_via_mailbox):
    _, bad, _ = explore(2, 1, 2, parities=1, order=order, self_via_mailbox=self_via_mailbox)
    assert bad is not None and "already holds step 2" in bad
