"""Oracle-side comparison of one KITTI sparse-decoder sample against ``oracle.kitti.sparse_forward``.

TEST INFRASTRUCTURE (see oracle/__init__.py): used by tests/ and by bench.py's ``cpu_baseline`` leg, never by
the product.

The parity statement (BASELINE.json north_star): floats within 1e-4 relative fp32 tolerance, active-pixel masks
bit-exact, ``total_ops`` exact.  A threshold test ``max|yh| > thresh`` (depth_decoder.py:308-309) evaluated on two
fp32 implementations that sum in different orders can legitimately disagree on a pixel whose margin is itself below
the float tolerance (a TIE).  This module therefore reports, per scale, the Hamming distance of every mask and,
when a mask differs, checks that each differing pixel of the threshold set S0 is such a tie in the ORACLE's own
numbers: ``| max_band|yh| - thresh | <= tie_tol * thresh``.  Outputs of the scales at and below the first differing
scale are then excluded from the float comparison (they legitimately evaluate a different active set); everything
coarser is still held to the float bar.  A differing pixel that is not a tie is a parity failure.
"""
import torch

MASK_NAMES = ("lowres_mask", "upconv0_mask", "upsample_mask", "upconv1_mask", "wavelet_mask")
BANDS = ("LH", "HL", "HH")


def rel_err(a, b):
    """max|a-b| / max(|b|, tiny) in float64: the 'relative fp32 tolerance' of the parity statement."""
    a = torch.as_tensor(a).detach().to("cpu", torch.float64)
    b = torch.as_tensor(b).detach().to("cpu", torch.float64)
    if a.numel() == 0:
        return 0.0
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def sample_of(out, b):
    """Sample b of a batched output dict as a batch-1 dict (tensors sliced, per-sample op count if present)."""
    res = {}
    for k, v in out.items():
        if torch.is_tensor(v):
            res[k] = v[b:b + 1]
    if "total_ops_per_sample" in out:
        res["total_ops"] = out["total_ops_per_sample"][b]
    elif "total_ops" in out and next(iter(res.values())).shape[0] == 1:
        res["total_ops"] = out["total_ops"]
    return res


def compare_kitti_sample(got, ref, thresh_ratio, float_tol=1e-4, tie_tol=1e-4):
    """got / ref: batch-1 output dicts of the sparse decoder (ref from oracle.kitti.sparse_forward).

    Returns a report dict:
      max_rel_err            worst rel_err over the float outputs that were compared
      mask_hamming           {scale: differing pixels summed over the five masks of that scale}
      wavelet_mask_hamming   {scale: differing pixels of ("wavelet_mask", scale)}
      total_ops_equal        bool (None if `got` carries no count)
      first_diff_scale       coarsest scale whose masks differ, or None
      ties_explained         True when every differing threshold pixel is a tie (vacuously True without differences)
      failures               list of strings; empty = parity holds
    """
    rep = {"max_rel_err": 0.0, "mask_hamming": {}, "wavelet_mask_hamming": {}, "total_ops_equal": None,
           "first_diff_scale": None, "ties_explained": True, "failures": []}
    for s in range(3, -1, -1):
        ham = 0
        for name in MASK_NAMES:
            g = got[(name, s)].detach().cpu().bool()
            r = ref[(name, s)].detach().cpu().bool()
            if g.shape != r.shape:
                rep["failures"].append("%s,%d shape %s vs %s" % (name, s, tuple(g.shape), tuple(r.shape)))
                continue
            d = int((g != r).sum())
            ham += d
            if name == "wavelet_mask":
                rep["wavelet_mask_hamming"][s] = d
        rep["mask_hamming"][s] = ham
        if ham and rep["first_diff_scale"] is None:
            rep["first_diff_scale"] = s
    fd = rep["first_diff_scale"]
    if fd is not None:
        # threshold set of the level that produced scale fd: S0 = max_band|yh(scale fd+1)| > (max-min)(LL entering it) * ratio
        i = fd + 1
        yl = ref[("wavelets", i - 1, "LL")].detach().cpu().double()
        thresh = float(yl.max() - yl.min()) * float(thresh_ratio)
        mag = torch.stack([ref[("wavelets", i, b)].detach().cpu().double().abs() for b in BANDS]).max(0)[0]
        flips = (got[("wavelet_mask", fd)].detach().cpu().bool() != ref[("wavelet_mask", fd)].detach().cpu().bool())
        flips_lo = flips[..., ::2, ::2]                       # wavelet_mask = nearest x2 of S0
        if int(flips_lo.sum()) * 4 != int(flips.sum()) or not bool(flips_lo.any()):
            rep["ties_explained"] = False
            rep["failures"].append("scale %d: mask differences are not 2x2 blocks of threshold pixels" % fd)
        else:
            margin = (mag[flips_lo] - thresh).abs() / max(thresh, 1e-30)
            worst = float(margin.max())
            rep["tie_margin"] = worst
            if worst > tie_tol:
                rep["ties_explained"] = False
                rep["failures"].append("scale %d: %d threshold pixels differ, worst margin %.3e > %.1e"
                                       % (fd, int(flips_lo.sum()), worst, tie_tol))
    for s in range(3, -1, -1):
        strict = fd is None or s > fd
        keys = [("wavelets", s, "LL")] if (strict or s == fd) else []
        if strict:
            keys += [("wavelets", s, b) for b in BANDS] + [("disp", s)]
        for k in keys:
            e = rel_err(got[k], ref[k])
            rep["max_rel_err"] = max(rep["max_rel_err"], e)
            if e > float_tol:
                rep["failures"].append("%s rel err %.3e > %.1e" % ("_".join(str(v) for v in k), e, float_tol))
    if "total_ops" in got:
        rep["total_ops_equal"] = int(got["total_ops"]) == int(ref["total_ops"])
        if fd is None and not rep["total_ops_equal"]:
            rep["failures"].append("total_ops %d != %d" % (int(got["total_ops"]), int(ref["total_ops"])))
    return rep


def merge_reports(reports):
    """Aggregate per-sample reports into the bench line's `parity` object."""
    agg = {"samples": len(reports), "max_rel_err": 0.0, "mask_hamming_per_scale": {str(s): 0 for s in range(4)},
           "wavelet_mask_hamming_per_scale": {str(s): 0 for s in range(4)}, "total_ops_equal": True,
           "samples_with_mask_differences": 0, "ties_explained": True, "failures": []}
    for r in reports:
        agg["max_rel_err"] = max(agg["max_rel_err"], r["max_rel_err"])
        for s in range(4):
            agg["mask_hamming_per_scale"][str(s)] += r["mask_hamming"].get(s, 0)
            agg["wavelet_mask_hamming_per_scale"][str(s)] += r["wavelet_mask_hamming"].get(s, 0)
        if r["first_diff_scale"] is not None:
            agg["samples_with_mask_differences"] += 1
        elif r["total_ops_equal"] is False:
            agg["total_ops_equal"] = False
        agg["ties_explained"] = agg["ties_explained"] and r["ties_explained"]
        agg["failures"] += r["failures"]
    agg["max_rel_err"] = float("%.3e" % agg["max_rel_err"])
    agg["failures"] = agg["failures"][:8]
    return agg
