"""Rebind the reference's module attributes to the GPU path (the drop-in hook).

The reference has no plugin ABI: callers reach the hot path through module globals and names
imported BY VALUE, so a replacement must be rebound at every importer:

    matching.MATCHER / MATCHER_LEFT_RIGHT / extract_inliers_outliers   (matching.py:48-73)
    database.MATCHER, database.extract_inliers_outliers,
    database.ransac_pnp_for_tracking_db                                 (database.py:4-6)
    loop_closure.MATCHER, loop_closure.ransac_pnp                       (loop_closure.py:6,12)
    ransac.triangulate_links, ransac.transformation_agreement,
    ransac.ransac_pnp_for_tracking_db, ransac.ransac_pnp                (ransac.py:4,28,70,116)
    triangulation.* , bundle.triangulate_last_frame (bundle.py:10),
    gtsam_utils.triangulate_last_frame (gtsam_utils.py:7), analysis.linear_least_squares_triangulation
"""
from __future__ import annotations

import sys

_REBINDS = {
    "final_project.algorithms.matching": ("MATCHER", "MATCHER_LEFT_RIGHT", "extract_inliers_outliers"),
    "final_project.backend.database.database": ("MATCHER", "extract_inliers_outliers", "ransac_pnp_for_tracking_db"),
    "final_project.backend.loop.loop_closure": ("MATCHER", "ransac_pnp"),
    "final_project.algorithms.ransac": ("triangulate_links", "transformation_agreement",
                                        "ransac_pnp_for_tracking_db", "ransac_pnp"),
    "final_project.algorithms.triangulation": ("linear_least_squares_triangulation", "triangulate_links",
                                               "triangulate_last_frame"),
    "final_project.backend.GTSam.bundle": ("triangulate_last_frame",),
    "final_project.backend.GTSam.gtsam_utils": ("triangulate_last_frame",),
    "final_project.analysis": ("linear_least_squares_triangulation",),
}


def replacements():
    from . import matching, ransac, triangulation
    if matching.MATCHER is None:
        matching.init_globals()
    return {
        "MATCHER": matching.MATCHER,
        "MATCHER_LEFT_RIGHT": matching.MATCHER_LEFT_RIGHT,
        "extract_inliers_outliers": matching.extract_inliers_outliers,
        "linear_least_squares_triangulation": triangulation.linear_least_squares_triangulation,
        "triangulate_links": triangulation.triangulate_links,
        "triangulate_last_frame": triangulation.triangulate_last_frame,
        "transformation_agreement": ransac.transformation_agreement,
        "ransac_pnp_for_tracking_db": ransac.ransac_pnp_for_tracking_db,
        "ransac_pnp": ransac.ransac_pnp,
    }


def _akaze_feature():
    """The detector of get_akaze_matcher_lr_matcher() (matching.py:19-21)."""
    import cv2
    return cv2.AKAZE_create(threshold=0.0008, nOctaves=4, nOctaveLayers=4)


def patch(modules=None, batched_db=False, batched_loop=False, rebind_feature=True, batched_gating=False,
          **pipeline_kw):
    """Rebind every already-imported reference module (or the given {name: module} mapping).
    Also copies the reference's cameras into slamfe.ransac so both sides score with the same
    K, M1, M2.  Returns {module_name: [rebound attribute names]}; `unpatch(token)` restores.

    The checked-in reference selects SIFT + L2 (matching.py:72; float32 128-d descriptors), which the
    Hamming-only GPU matchers cannot serve: with rebind_feature=True (default) `matching.FEATURE` is rebound
    to the AKAZE detector of the reference's own get_akaze_matcher_lr_matcher() (matching.py:19-21) whenever
    the current FEATURE does not produce 8-bit descriptors; with rebind_feature=False such a configuration is
    refused with a TypeError instead of failing later inside cv2.

    batched_db=True additionally replaces `database.create_db` (database.py:30, called by
    `database.run`, :92-98) with the batched whole-sequence builder of slamfe.database: frames are
    still read and described on the CPU exactly as the reference does, everything else of the loop
    runs as one GPU pipeline (hypotheses from the GPU generator).

    batched_loop=True replaces `loop_closure.check_candidate_match` / `consensus_matches`
    (loop_closure.py:405-436, :572-599) with slamfe.loop's: all candidates of a keyframe are verified
    in one device-resident batch (match + 888-hypothesis RANSAC-PnP each).

    batched_gating=True replaces `loop_closure.get_good_candidates` (loop_closure.py:199-228) with
    slamfe.loop.get_good_candidates_typed bound to the module's own `cov_dijkstra_graph` /
    `relative_covariance_dict` (read at call time): every earlier keyframe is gated in one launch."""
    from . import ransac
    rep = replacements()
    mods = modules if modules is not None else sys.modules
    done, saved = {}, []
    ref_ransac = mods.get("final_project.algorithms.ransac")
    if ref_ransac is not None and hasattr(ref_ransac, "K"):
        ransac.set_cameras(ref_ransac.K, ref_ransac.M1, ref_ransac.M2)
    ref_matching = mods.get("final_project.algorithms.matching")
    feature = getattr(ref_matching, "FEATURE", None) if ref_matching is not None else None
    if feature is not None and hasattr(feature, "descriptorType") and feature.descriptorType() != 0:   # CV_8U == 0
        if not rebind_feature:
            raise TypeError("the reference's matching.FEATURE produces non-binary descriptors (SIFT/L2, matching.py:72); "
                            "slamfe implements the AKAZE/Hamming configuration (matching.py:19-24)")
        saved.append((ref_matching, "FEATURE", feature))
        ref_matching.FEATURE = _akaze_feature()
        done.setdefault("final_project.algorithms.matching", []).append("FEATURE")
    for mod_name, attrs in _REBINDS.items():
        mod = mods.get(mod_name)
        if mod is None:
            continue
        for a in attrs:
            if hasattr(mod, a):
                saved.append((mod, a, getattr(mod, a)))
                setattr(mod, a, rep[a])
                done.setdefault(mod_name, []).append(a)
    lcm = mods.get("final_project.backend.loop.loop_closure")
    if batched_loop and lcm is not None:
        from . import loop as sloop
        for a, fn in (("check_candidate_match", sloop.check_candidate_match), ("consensus_matches", sloop.consensus_matches)):
            if hasattr(lcm, a):
                saved.append((lcm, a, getattr(lcm, a)))
                setattr(lcm, a, fn)
                done.setdefault("final_project.backend.loop.loop_closure", []).append(a)
    if batched_gating and lcm is not None and hasattr(lcm, "get_good_candidates"):
        from . import loop as sloop

        def get_good_candidates(c_n_index, marginals, result, index_list, _m=lcm):
            sym = getattr(getattr(_m, "gtsam", None), "symbol", None)
            return sloop.get_good_candidates_typed(c_n_index, marginals, result, index_list, _m.cov_dijkstra_graph,
                                                   _m.relative_covariance_dict, symbol=sym)

        saved.append((lcm, "get_good_candidates", lcm.get_good_candidates))
        lcm.get_good_candidates = get_good_candidates
        done.setdefault("final_project.backend.loop.loop_closure", []).append("get_good_candidates")
    if batched_db:
        dbm = mods.get("final_project.backend.database.database")
        if dbm is not None and hasattr(dbm, "create_db"):
            from . import database as sdb
            saved.append((dbm, "create_db", dbm.create_db))
            dbm.create_db = sdb.make_reference_create_db(dbm, matching_module=mods.get(
                "final_project.algorithms.matching"), **pipeline_kw)
            done.setdefault("final_project.backend.database.database", []).append("create_db")
    done["_saved"] = saved
    return done


def unpatch(token):
    for mod, a, old in token.get("_saved", []):
        setattr(mod, a, old)
