"""Winner selection of the exact-pass kernel (csrc/merge.cuh: select_winners, select_winners_cluster), restated in numpy.

The streaming kernel's CTAs publish sorted lists of at most kc distinct keys.  The selection prunes with
    T0 = max over FULL lists of their kc-th key         (that one list alone holds kc keys >= T0)
    T1 = kc-th largest of the first P = ceil(kc / lists) keys of every list
and keeps the keys >= max(T0, T1); the cluster-wide variant lets every CTA scan a slice of the published regions, all-gathers the
survivors and has each CTA rank the survivors it gathered itself by counting the greater ones.  Checked here: the threshold never
cuts into the true top-kc, the per-owner ranks form a permutation, and the result equals a plain sort — for ragged lists, empty
lists, fewer than kc keys in total."""
import numpy as np
import pytest

REGIONS, CSIZE = 16, 8


def _publish(rng, lists, kc, fill):
    """lists sorted lists (descending) of distinct uint64 keys; list l goes to region l % 16 in arrival order."""
    total = int(sum(fill))
    keys = rng.choice(np.arange(1, 50 * max(total, 1) + 2, dtype=np.uint64), size=total, replace=False)
    out, pos = [], 0
    for l in range(lists):
        out.append(np.sort(keys[pos:pos + fill[l]])[::-1])
        pos += fill[l]
    return out


def _select(published, kc):
    lists = len(published)
    P = -(-kc // lists)
    heads = np.concatenate([np.pad(l[:P], (0, P - len(l[:P]))) for l in published]) if lists else np.zeros(0, dtype=np.uint64)
    tails = np.array([l[kc - 1] if len(l) >= kc else 0 for l in published], dtype=np.uint64)
    t0 = tails.max() if lists else 0
    nz = np.sort(heads[heads != 0])[::-1]
    t1 = nz[kc - 1] if len(nz) >= kc else 0
    T = max(t0, t1)
    # regions, then the cluster-wide split: CTA c owns regions [2c, 2c + 2)
    regions = [[] for _ in range(REGIONS)]
    for l, ks in enumerate(published):
        regions[l % REGIONS].extend(ks.tolist())
    owned = [[k for r in range(c * REGIONS // CSIZE, (c + 1) * REGIONS // CSIZE) for k in regions[r] if k >= T and k != 0] for c in range(CSIZE)]
    everything = np.array([k for o in owned for k in o], dtype=np.uint64)
    out = np.zeros(kc, dtype=np.uint64)
    written = np.zeros(kc, dtype=int)
    for c in range(CSIZE):                                          # each CTA ranks the survivors it gathered itself
        for k in owned[c]:
            rank = int((everything > k).sum())
            if rank < kc:
                out[rank] = k
                written[rank] += 1
    return T, everything, out, written


@pytest.mark.parametrize("kc,lists,mode", [(32, 148, "full"), (128, 148, "full"), (256, 148, "full"), (32, 48, "short"), (64, 148, "ragged"),
                                           (128, 71, "ragged"), (32, 148, "sparse"), (256, 5, "short"), (32, 1, "full")])
def test_selection_equals_a_plain_sort(kc, lists, mode):
    rng = np.random.default_rng(kc * 1000 + lists)
    for trial in range(5):
        if mode == "full":
            fill = [kc] * lists
        elif mode == "short":
            fill = rng.integers(0, max(2, kc // 3), size=lists).tolist()
        elif mode == "ragged":
            fill = rng.integers(0, kc + 1, size=lists).tolist()
        else:                                                       # a handful of keys in total: fewer than kc winners exist
            fill = (rng.random(lists) < 0.05).astype(int).tolist()
        published = _publish(rng, lists, kc, fill)
        allk = np.sort(np.concatenate(published))[::-1] if sum(fill) else np.zeros(0, dtype=np.uint64)
        T, survivors, out, written = _select(published, kc)
        want = np.zeros(kc, dtype=np.uint64)
        want[:min(kc, len(allk))] = allk[:kc]
        if len(allk) >= kc:
            assert T <= allk[kc - 1]                                # the threshold is a lower bound of the kc-th best key
        m = min(kc, len(survivors))
        assert np.all(written[:m] == 1) and np.all(written[m:] == 0)    # ranks are a permutation: every slot written exactly once
        assert np.array_equal(out, want)
