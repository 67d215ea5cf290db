#!/usr/bin/env python3
"""Camera frames -> poses on one B200, device-resident (DeviceLocalizer.run_frames), the production tick of
visual_landmark_matcher.py:293-380 for batches of frames: ORB(500) extraction, crossCheck matching against the <= 5
candidate keyframes nearest along the route (here: the true keyframe and its four neighbours), PnP-RANSAC on every
candidate with >= 10 matches, gates, best candidate.  The library is TAUGHT FROM IMAGES on the same GPU (ORB + keypoint
lifting per teach frame).  One JSON object: frames/s, the share of ORB, PnP problems per frame, and the check that
every query frame localises to the teach frame it shows."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--keyframes', type=int, default=400)
    ap.add_argument('--batch', type=int, default=128)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--depth', type=int, default=2, help='engines of the PipelinedFrameLocalizer')
    args = ap.parse_args()
    import torch
    import nclt_slam_project_b200  # noqa
    from nclt_slam_project_b200 import synth
    from nclt_slam_project_b200.orb import ORB
    from nclt_slam_project_b200.pipeline import DeviceLocalizer, LocalizeParams
    from nclt_slam_project_b200.recorder import lift_keypoints
    K, B = args.keyframes, args.batch
    uniq = np.stack([synth.make_camera_frame(1000 + s) for s in range(K)])
    rng = np.random.default_rng(4)
    depth = (5000 + 3 * rng.standard_normal((480, 640))).astype(np.uint16)
    # teach: ORB + lifting for every keyframe (batched)
    orb = ORB(max_frames=64)
    descs, pts3 = [], []
    t0 = time.perf_counter()
    for i in range(0, K, 64):
        kp, desc, n = orb.detect_and_compute_batch(uniq[i:i + 64])
        F = len(n)
        res = lift_keypoints(kp[:, :, :2].copy(), np.broadcast_to(depth, (F, 480, 640)).copy(), n_kpts=n)
        for f in range(F):
            keep, p3 = res[f]
            descs.append(desc[f][keep])
            pts3.append(p3)
    teach_s = time.perf_counter() - t0
    dl = DeviceLocalizer((descs, pts3), params=LocalizeParams(mode=1))
    order = rng.permutation(K)[:B] if B <= K else rng.integers(0, K, B)
    q = torch.from_numpy(uniq[order]).to(dl.device)
    cand_np = ((order[:, None] + np.array([-2, -1, 0, 1, 2])[None, :]) % K).astype(np.int32)     # matcher:293-302
    cand = torch.from_numpy(cand_np).to(dl.device)
    out = dl.run_frames(q, cand)
    torch.cuda.synchronize()
    slot = out['best_cand'].cpu().numpy()
    ok = float((cand_np[np.arange(B), np.maximum(slot, 0)] == order).mean())
    n_problems = out['n_problems']
    mean_inl = float(out['n_inliers'].float().mean())
    for _ in range(2):
        dl.run_frames(q, cand)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dl.run_frames(q, cand)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / args.steps
    # two engines alternate (pipeline.PipelinedFrameLocalizer, one host thread): the latency-bound PnP tail of one batch
    # (EPnP rounds, LM finish: a few warps per SM) runs beside the ORB kernels of the next batch
    from nclt_slam_project_b200.pipeline import PipelinedFrameLocalizer
    D = args.depth
    pfl = PipelinedFrameLocalizer((descs, pts3), depth=D)
    outs = [None] * D
    for i in range(2 * D):
        outs[i % D] = pfl.submit(q, cand)[1]
    pfl.synchronize()
    t0 = time.perf_counter()
    for i in range(2 * args.steps):
        outs[i % D] = pfl.submit(q, cand)[1]
    pfl.synchronize()
    dt2 = (time.perf_counter() - t0) / (2 * args.steps)
    assert pfl.overflow() == 0, 'PnP capacity overflow in the asynchronous tick'
    for o2 in outs:
        for k in ('best_cand', 'n_inliers', 'rvec', 'tvec'):
            assert torch.equal(torch.nan_to_num(o2[k].double()), torch.nan_to_num(out[k].double())), k
    # ORB alone on the same frames
    kp, desc, n = dl._orb_out
    from nclt_slam_project_b200._lib import lib as L
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dl.ctx.check(L.nclt_orb_detect_and_compute_dev(dl.ctx.h, dl._orb._h, q.data_ptr(), 1, B, kp.data_ptr(), desc.data_ptr(), n.data_ptr()))
    torch.cuda.synchronize()
    dt_orb = (time.perf_counter() - t0) / args.steps
    # one profiled step: device time per kernel family (CUDA events around the launches, nclt_ctx_profile_read_tags)
    dl.ctx.profile(True)
    dl.ctx.profile_read_tags()
    dl.run_frames(q, cand, sync_count=False)
    fam = {k: {'ms': round(v[0], 4), 'launches': v[1]} for k, v in dl.ctx.profile_read_tags().items() if v[1]}
    dl.ctx.profile(False)
    print(json.dumps({'kernel_families_ms_per_step': fam, 'workload': f'{B} gray 640x480 frames per step vs a {K}-keyframe library taught from images '
                                  f'({int(np.mean([len(d) for d in descs]))} landmarks per keyframe), crossCheck against 5 candidate keyframes + PnP-RANSAC',
                      'pnp_problems_per_frame': n_problems / B, 'frames_per_s': B / dt, 'ms_per_step': dt * 1e3,
                      'frames_per_s_two_contexts': B / dt2, 'ms_per_step_two_contexts': dt2 * 1e3, 'pipelined_engines': D, 'orb_share_of_step': dt_orb / dt,
                      'orb_frames_per_s': B / dt_orb, 'localised_to_own_keyframe': ok,
                      'mean_inliers': mean_inl,
                      'teach_frames_per_s_host_driven': K / teach_s}))


if __name__ == '__main__':
    main()
