"""Synthetic DBoW2-style feature vectors for the SearchByBoW tests: a 'vocabulary node' is (octave, coarse cell of the
flow-compensated position), so that corresponding features of the two frames mostly share a node, the way descriptors
of the same scene point fall into the same vocabulary word."""
import numpy as np

import oracle_lib as O
from pilotguru_b200 import synth


def feats(t, w=640, h=480, nf=500):
    orc = O.OrbOracle(nf, 1.2, 8, 20, 7)
    return orc.extract(synth.frame(t, w=w, h=h))


def featvec(kps, shift, cell, rng=None, drop=0.0):
    """{node id: [feature indices]}; node = octave * 4096 + cell_x * 64 + cell_y of (position - shift).  High octaves
    use one node per octave (large nodes).  With rng: a fraction `drop` of the features is left out and the lists of
    every third node are shuffled (the reference walks them in vector order, whatever it is)."""
    fv = {}
    for i in range(len(kps)):
        if rng is not None and rng.uniform() < drop:
            continue
        o = int(kps["octave"][i])
        cx = int((kps["x"][i] - shift[0]) // cell) + 1
        cy = int((kps["y"][i] - shift[1]) // cell) + 1
        node = o * 4096 + (0 if o >= 5 else cx * 64 + cy)
        fv.setdefault(node, []).append(i)
    if rng is not None:
        for j, k in enumerate(sorted(fv)):
            if j % 3 == 0:
                rng.shuffle(fv[k])
    return fv


def problem(t_kf, t_f, seed=0, cell=80, w=640, h=480):
    rng = np.random.default_rng(seed)
    (kk, kd), (fk, fd) = feats(t_kf, w, h), feats(t_f, w, h)
    shift = np.sum([synth.flow(t, w=w, h=h) for t in range(t_kf + 1, t_f + 1)], axis=0) if t_f > t_kf else np.zeros(2)
    kfv = featvec(kk, (0.0, 0.0), cell, rng, drop=0.05)
    ffv = featvec(fk, shift, cell, rng, drop=0.05)
    has = (rng.uniform(size=len(kk)) > 0.15).astype(np.uint8)
    return dict(kf_desc=kd, kf_angle=kk["angle"].astype(np.float32), kf_has=has, kf_fv=kfv,
                f_desc=fd, f_angle=fk["angle"].astype(np.float32), f_fv=ffv)


def python_search_by_bow(P, nnratio, check_ori):
    """Independent restatement with dicts and numpy popcounts (set intersection instead of the merge loop)."""
    kd, fd = P["kf_desc"], P["f_desc"]
    match = np.full(len(fd), -1, np.int64)
    hist = [[] for _ in range(30)]
    n = 0
    for node in sorted(set(P["kf_fv"]) & set(P["f_fv"])):
        for ik in P["kf_fv"][node]:
            if not P["kf_has"][ik]:
                continue
            cands = [(int(np.unpackbits(kd[ik] ^ fd[jf]).sum()), pos, jf) for pos, jf in enumerate(P["f_fv"][node]) if match[jf] < 0]
            if not cands:
                continue
            cands.sort()
            b1 = cands[0][0]; b2 = cands[1][0] if len(cands) > 1 else 256
            if b1 <= 50 and np.float32(b1) < np.float32(nnratio) * np.float32(b2):
                jf = cands[0][2]
                match[jf] = ik; n += 1
                if check_ori:
                    rot = np.float32(P["kf_angle"][ik]) - np.float32(P["f_angle"][jf])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    v = np.float32(rot * np.float32(1.0 / 30))
                    b = int(np.floor(v + np.float32(0.5))) if v >= 0 else int(np.ceil(v - np.float32(0.5)))
                    hist[0 if b == 30 else b].append(jf)
    if check_ori:
        sizes = [len(x) for x in hist]
        m1 = m2 = m3 = 0; i1 = i2 = i3 = -1
        for i, s in enumerate(sizes):
            if s > m1:
                m3, m2, m1 = m2, m1, s; i3, i2, i1 = i2, i1, i
            elif s > m2:
                m3, m2 = m2, s; i3, i2 = i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        for i in range(30):
            if i in (i1, i2, i3):
                continue
            for jf in hist[i]:
                match[jf] = -1; n -= 1
    return n, match
