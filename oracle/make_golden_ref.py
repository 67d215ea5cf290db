"""TEST INFRASTRUCTURE - golden vectors produced by the REFERENCE'S OWN MODULES, imported
unmodified from /root/reference under ROS stubs (build container only).

  golden_map      tf_wall_clock_relay.TFRelay.depth_cb -> teach_run_depth_mapper.TeachDepthMapper.cb
                  -> save(): point clouds, log-odds grid, PGM/YAML bytes on seeded synthetic depth
  golden_selftest the candidate loop of checkpoint_a_selftest.run_matcher_self (lines 62-103) with
                  the module's own imported constants, fed synthetic descriptors
"""
import io
import os
import tempfile

import numpy as np

from . import ros_stubs
from nclt_slam_project_b200 import synth

# a small grid keeps the fixture small; the CLI defaults are exercised by the format test
MAP_CFG = dict(origin_x=-12.0, origin_y=-9.0, width_m=40.0, height_m=30.0, res=0.1)


def _fake_image(depth):
    from sensor_msgs.msg import Image
    m = Image()
    if depth.dtype == np.uint16:
        m.encoding = '16UC1'
    else:
        m.encoding = '32FC1'
    m.height, m.width = depth.shape
    m.data = depth.tobytes()
    return m


def golden_map(out_dir, n_frames=24):
    mods = ros_stubs.import_reference()
    relay_mod, mapper_mod = mods['tf_wall_clock_relay'], mods['teach_run_depth_mapper']
    relay = relay_mod.TFRelay(use_gt=True)
    tmp = tempfile.mkdtemp()
    prefix = os.path.join(tmp, 'teach_map')
    mapper = mapper_mod.TeachDepthMapper(prefix, MAP_CFG['origin_x'], MAP_CFG['origin_y'], MAP_CFG['width_m'],
                                         MAP_CFG['height_m'], MAP_CFG['res'])
    rng = np.random.default_rng(3)
    poses, depths, clouds, cloud_n, grids_after = [], [], [], [], []
    x, y, yaw = -8.0, -5.0, 0.3
    for f in range(n_frames):
        x += 0.35 * np.cos(yaw)
        y += 0.35 * np.sin(yaw)
        yaw += rng.uniform(-0.05, 0.25)
        if f == 5:
            pose = (-30.0, 0.0, 0.0)            # sensor cell outside the grid: frame dropped
        else:
            pose = (x, y, yaw)
        depth = synth.make_depth_frame(100 + f, pose, cyl_density=0.05)
        if f == 7:
            depth[:] = 0.0                      # empty cloud
        if f == 9:
            depth = (np.nan_to_num(depth, nan=0.0, posinf=0.0) * 1000.0).clip(0, 65535).astype(np.uint16)
        tf = synth.camera_link_transform(*pose)
        relay.depth_cb(_fake_image(depth))
        pc = relay.pc_pub.sent[-1]
        pts = np.frombuffer(pc.data, dtype=np.float32).reshape(-1, 3)
        assert pc.point_step == 12 and pc.width == len(pts) and pc.header.frame_id == 'camera_link'
        mapper.tf_buf.current = tf
        mapper.cb(pc)
        poses.append(tf)
        if depth.dtype == np.uint16:
            depth_u16 = depth.copy()
        depths.append(depth.astype(np.float32))
        clouds.append(pts.copy())
        cloud_n.append(len(pts))
        if f in (0, 3, n_frames - 1):
            grids_after.append(mapper.grid.copy())
    mapper.save()
    pgm = open(prefix + '.pgm', 'rb').read()
    yml = open(prefix + '.yaml', 'rb').read().replace(tmp.encode(), b'<TMP>')
    nmax = max(cloud_n)
    cl = np.zeros((n_frames, nmax, 3), dtype=np.float32)
    for f, c in enumerate(clouds):
        cl[f, :len(c)] = c
    # depth frames as float32 metres; frame 9 was fed as 16UC1 millimetres (depth_u16)
    np.savez_compressed(
        os.path.join(out_dir, 'map_golden.npz'),
        cfg=np.array([MAP_CFG['origin_x'], MAP_CFG['origin_y'], MAP_CFG['width_m'], MAP_CFG['height_m'], MAP_CFG['res']]),
        tf=np.array(poses), depth=np.stack(depths).astype(np.float32), u16_frames=np.array([9], dtype=np.int32), depth_u16=depth_u16,
        cloud=cl, cloud_n=np.array(cloud_n, dtype=np.int32),
        grid_snap_frames=np.array([0, 3, n_frames - 1], dtype=np.int32), grid_snaps=np.stack(grids_after),
        grid_final=mapper.grid.copy(), pgm=np.frombuffer(pgm, dtype=np.uint8), yaml=np.frombuffer(yml, dtype=np.uint8),
        frames_integrated=np.int64(mapper.frames_integrated), total_points=np.int64(mapper.total_points_integrated),
        skipped_empty=np.int64(mapper.frames_skipped_empty))
    print('map_golden.npz frames_integrated', mapper.frames_integrated, 'pts', mapper.total_points_integrated,
          'pgm bytes', len(pgm), 'clouds', cloud_n[:6])


def golden_selftest(out_dir):
    """Run the reference selftest's candidate loop (checkpoint_a_selftest.py:62-103) on synthetic
    descriptors. run_matcher_self() itself starts from JPEG files (cv2.imread + ORB), which are not
    shipped, so its loop body is executed here on the module's own imported names."""
    mods = ros_stubs.import_reference()
    st = mods['checkpoint_a_selftest']
    import cv2
    data = synth.make_library(77, n_kf=10, n_desc=400, ragged=True)
    rows = []
    descs, pts2, cands = [], [], []
    for seed in range(6):
        f = synth.make_frame(data, 7700 + seed, n_desc=500, n_planted=220)
        cand_idx = [f['k_star'], (f['k_star'] + 1) % 10, (f['k_star'] + 5) % 10]
        desc_curr, pts_curr_2d = f['desc'], f['pts2d']
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
        best = None
        per = []
        for li in cand_idx:
            lm_t = data['landmarks'][li]
            desc_t = lm_t['descriptors']
            rec = [0, 0, 0, 0.0, 0, 0, 0, 0, 0, 0]
            per.append(rec)
            if desc_t is None or len(desc_t) < st.MIN_MATCHES:
                continue
            knn = bf.knnMatch(desc_curr, desc_t, k=2)
            good = [m for m, n in knn if (m.distance < st.LOWE_RATIO * n.distance)]
            rec[0] = len(good)
            if len(good) < st.MIN_MATCHES:
                continue
            obj_pts = np.array([lm_t['keypoints_3d_cam'][m.trainIdx] for m in good], dtype=np.float32)
            img_pts = np.array([pts_curr_2d[m.queryIdx] for m in good], dtype=np.float32)
            ok, rvec, tvec, inliers = cv2.solvePnPRansac(
                obj_pts, img_pts, st.K, st.DIST, iterationsCount=st.RANSAC_ITERATIONS,
                reprojectionError=st.RANSAC_REPROJ_PX, flags=cv2.SOLVEPNP_ITERATIVE)
            if not ok or inliers is None or len(inliers) < st.MIN_INLIERS:
                continue
            proj, _ = cv2.projectPoints(obj_pts[inliers[:, 0]], rvec, tvec, st.K, st.DIST)
            err = float(np.linalg.norm(proj.reshape(-1, 2) - img_pts[inliers[:, 0]], axis=1).mean())
            rec[1:] = [1, len(inliers), err] + rvec.ravel().tolist() + tvec.ravel().tolist()
            if err > st.REPROJ_MAX_PX:
                continue
            if best is None or len(inliers) > best[1]:
                best = (cand_idx.index(li), len(inliers))
        rows.append((best[0] if best else -1, best[1] if best else 0, per))
        descs.append(desc_curr)
        pts2.append(pts_curr_2d)
        cands.append(cand_idx)
    np.savez_compressed(
        os.path.join(out_dir, 'selftest_golden.npz'),
        lib_seed=np.int64(77), desc=np.stack(descs), pts2d=np.stack(pts2), cand=np.array(cands, dtype=np.int32),
        best_slot=np.array([r[0] for r in rows], dtype=np.int32), best_inl=np.array([r[1] for r in rows], dtype=np.int32),
        items=np.array([r[2] for r in rows], dtype=np.float64))
    print('selftest_golden.npz best slots', [r[0] for r in rows], 'inliers', [r[1] for r in rows])


def golden_selftest_short(out_dir):
    """The same loop body (checkpoint_a_selftest.py:62-103) on a library with 2-, 5-, 9- and 10-row keyframes whose
    first row attracts dozens of ratio matches from the frame (many-to-one): the reference skips the three
    keyframes with fewer than MIN_MATCHES rows BEFORE matching (selftest:64-65) although each of them would
    collect >= MIN_MATCHES `good` matches; the 10-row keyframe is matched and goes to solvePnPRansac."""
    mods = ros_stubs.import_reference()
    st = mods['checkpoint_a_selftest']
    import cv2
    data = synth.make_library(78, n_kf=8, n_desc=300, ragged=True)
    rng = np.random.default_rng(7801)
    short = {1: 2, 2: 5, 3: 9, 4: 10}
    for k, n in short.items():
        lm = data['landmarks'][k]
        for key in ('descriptors', 'keypoints_2d', 'keypoints_3d_cam'):
            lm[key] = np.ascontiguousarray(lm[key][:n])
        lm['n_features'] = n
    rows, descs, pts2, cands, would = [], [], [], [], []
    for seed in range(4):
        f = synth.make_frame(data, 7800 + seed, k_star=0 if seed < 2 else 6, n_desc=500, n_planted=200)
        desc_curr, pts_curr_2d = f['desc'].copy(), f['pts2d']
        # 30 frame rows per short keyframe = that keyframe's row 0 with a few flipped bits
        for j, k in enumerate(short):
            base = data['landmarks'][k]['descriptors'][0]
            for i in range(30):
                d = base.copy()
                for bit in rng.choice(256, size=int(rng.integers(0, 9)), replace=False):
                    d[bit >> 3] ^= np.uint8(1 << (bit & 7))
                desc_curr[300 + 30 * j + i] = d
        cand_idx = [1, 2, f['k_star'], 3, 4]
        bf = cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=False)
        best = None
        per = []
        w = []
        for li in cand_idx:
            lm_t = data['landmarks'][li]
            desc_t = lm_t['descriptors']
            rec = [0, 0, 0, 0.0, 0, 0, 0, 0, 0, 0]
            per.append(rec)
            knn_all = bf.knnMatch(desc_curr, desc_t, k=2)
            w.append(len([1 for m, n in knn_all if (m.distance < st.LOWE_RATIO * n.distance)]))
            if desc_t is None or len(desc_t) < st.MIN_MATCHES:
                continue
            knn = bf.knnMatch(desc_curr, desc_t, k=2)
            good = [m for m, n in knn if (m.distance < st.LOWE_RATIO * n.distance)]
            rec[0] = len(good)
            if len(good) < st.MIN_MATCHES:
                continue
            obj_pts = np.array([lm_t['keypoints_3d_cam'][m.trainIdx] for m in good], dtype=np.float32)
            img_pts = np.array([pts_curr_2d[m.queryIdx] for m in good], dtype=np.float32)
            ok, rvec, tvec, inliers = cv2.solvePnPRansac(
                obj_pts, img_pts, st.K, st.DIST, iterationsCount=st.RANSAC_ITERATIONS,
                reprojectionError=st.RANSAC_REPROJ_PX, flags=cv2.SOLVEPNP_ITERATIVE)
            if not ok or inliers is None or len(inliers) < st.MIN_INLIERS:
                continue
            proj, _ = cv2.projectPoints(obj_pts[inliers[:, 0]], rvec, tvec, st.K, st.DIST)
            err = float(np.linalg.norm(proj.reshape(-1, 2) - img_pts[inliers[:, 0]], axis=1).mean())
            rec[1:] = [1, len(inliers), err] + rvec.ravel().tolist() + tvec.ravel().tolist()
            if err > st.REPROJ_MAX_PX:
                continue
            if best is None or len(inliers) > best[1]:
                best = (cand_idx.index(li), len(inliers))
        rows.append((best[0] if best else -1, best[1] if best else 0, per))
        descs.append(desc_curr)
        pts2.append(pts_curr_2d)
        cands.append(cand_idx)
        would.append(w)
    would = np.array(would)
    assert (would[:, [0, 1, 3]] >= st.MIN_MATCHES).all(), would      # the skipped keyframes WOULD pass the count gate
    lms = data['landmarks']
    np.savez_compressed(
        os.path.join(out_dir, 'selftest_short_golden.npz'),
        counts=np.array([len(lm['descriptors']) for lm in lms], dtype=np.int32),
        lib_desc=np.concatenate([lm['descriptors'] for lm in lms]),
        lib_p3d=np.concatenate([lm['keypoints_3d_cam'] for lm in lms]),
        desc=np.stack(descs), pts2d=np.stack(pts2), cand=np.array(cands, dtype=np.int32),
        would_match=would.astype(np.int32),
        best_slot=np.array([r[0] for r in rows], dtype=np.int32), best_inl=np.array([r[1] for r in rows], dtype=np.int32),
        items=np.array([r[2] for r in rows], dtype=np.float64))
    print('selftest_short_golden.npz best slots', [r[0] for r in rows], 'inliers', [r[1] for r in rows],
          'ratio matches the skipped keyframes would have had', would.tolist(), 'items nmatch',
          [[int(x[0]) for x in r[2]] for r in rows])


def golden_tick(out_dir):
    """The production node: VisualLandmarkMatcher._tick (visual_landmark_matcher.py:281-433), run
    unmodified under ROS stubs with a stand-in ORB that returns synthetic keypoints/descriptors.
    ACCUM_ENABLE is switched off for the run (the accumulation branch is dead in production,
    matcher:78-89, and would mutate the landmark set between ticks)."""
    import pickle
    mods = ros_stubs.import_reference()
    vm = mods['visual_landmark_matcher']
    vm.ACCUM_ENABLE = False
    data = synth.make_library(55, n_kf=40, n_desc=300, ragged=True, route_len_m=80.0)
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'south_landmarks.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(data, f)
    csv = os.path.join(tmp, 'log', 'anchor_matches.csv')
    node = vm.VisualLandmarkMatcher(pkl, csv)
    node.last_rgb = np.zeros((480, 640, 3), dtype=np.uint8)
    node.last_depth = np.full((480, 640), 2000, dtype=np.uint16)

    class _Kp:
        def __init__(self, pt):
            self.pt = (float(pt[0]), float(pt[1]))

    class _Orb:
        def __init__(self):
            self.next = None

        def detectAndCompute(self, gray, mask):
            d, p = self.next
            return [_Kp(q) for q in p], d

    node.orb = _Orb()
    rng = np.random.default_rng(9)
    ticks = []
    kinds = ['ok', 'ok', 'far', 'few', 'random', 'shifted', 'ok', 'ok', 'ok', 'reverse']
    for i, kind in enumerate(kinds):
        k = int(rng.integers(3, 37))
        fr = synth.make_frame(data, 5500 + i, k_star=k, n_desc=500, n_planted=(0 if kind == 'random' else 200))
        lm = data['landmarks'][k]
        bx, by = lm['pose'][0] - 0.35 + rng.normal(0, 0.3), lm['pose'][1] + rng.normal(0, 0.3)
        yaw = rng.normal(0, 0.1)
        if kind == 'far':
            by += 40.0
        if kind == 'shifted':
            bx += 6.5
        if kind == 'reverse':
            yaw += np.pi
        desc, pts = fr['desc'], fr['pts2d']
        if kind == 'few':
            desc, pts = desc[:6], pts[:6]
        base_pose = (float(bx), float(by), 0.0, 0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2)))
        node.orb.next = (desc, pts)
        node._read_pose = (lambda bp=base_pose: bp)
        n_before = len(node.anchor_pub.sent)
        node._tick()
        line = open(csv).read().strip().split('\n')[-1].split(',')
        rec = {'kind': kind, 'base_pose': base_pose, 'csv': line[1:], 'published': len(node.anchor_pub.sent) > n_before}
        if rec['published']:
            m = node.anchor_pub.sent[-1]
            p, o = m.pose.pose.position, m.pose.pose.orientation
            rec['anchor'] = [p.x, p.y, p.z, o.x, o.y, o.z, o.w]
            rec['cov'] = list(m.pose.covariance)
        ticks.append((rec, desc, pts))
    header = open(csv).read().split('\n')[0]
    np.savez_compressed(
        os.path.join(out_dir, 'tick_golden.npz'), lib_seed=np.int64(55), header=np.array(header),
        kinds=np.array([t[0]['kind'] for t in ticks]), base_pose=np.array([t[0]['base_pose'] for t in ticks]),
        csv=np.array([','.join(t[0]['csv']) for t in ticks]), published=np.array([t[0]['published'] for t in ticks]),
        anchor=np.array([t[0].get('anchor', [0] * 7) for t in ticks], dtype=np.float64),
        cov=np.array([t[0].get('cov', [0] * 36) for t in ticks], dtype=np.float64),
        n_desc=np.array([len(t[1]) for t in ticks], dtype=np.int32),
        desc=np.stack([np.pad(t[1], ((0, 500 - len(t[1])), (0, 0))) for t in ticks]),
        pts2d=np.stack([np.pad(t[2], ((0, 500 - len(t[2])), (0, 0))) for t in ticks]))
    print('tick_golden.npz', [t[0]['csv'][-1] for t in ticks])


def golden_accum(out_dir):
    """The production node WITH continuous accumulation (visual_landmark_matcher.py:85 ACCUM_ENABLE = True, the shipped
    default; _maybe_accumulate :434-500): ticks far from every teach landmark after the silence period append the
    current frame as a new landmark; later ticks near it localise against the ACCUMULATED landmark.  The SIGTERM
    handler's augmented pickle (:192-202) is produced by invoking the node's own registered handler."""
    import pickle
    import signal
    mods = ros_stubs.import_reference()
    vm = mods['visual_landmark_matcher']
    vm.ACCUM_ENABLE = True
    data = synth.make_library(56, n_kf=20, n_desc=300, ragged=True, route_len_m=40.0)
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'south_landmarks.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(data, f)
    csv = os.path.join(tmp, 'log', 'anchor_matches.csv')
    node = vm.VisualLandmarkMatcher(pkl, csv)
    node.last_rgb = np.zeros((480, 640, 3), dtype=np.uint8)

    class _Kp:
        def __init__(self, pt):
            self.pt = (float(pt[0]), float(pt[1]))

    class _Orb:
        next = None

        def detectAndCompute(self, gray, mask):
            d, p = self.next
            return [_Kp(q) for q in p], d

    node.orb = _Orb()
    rng = np.random.default_rng(19)
    # a new place 30 m off the route: every frame shows the same scene (planted from a hidden "scene" keyframe with
    # known 3-D points, seen from slightly different poses), depth consistent with those points
    scene = synth.make_library(5601, n_kf=1, n_desc=400)
    ticks = []
    plan = [('near_route', 0.0), ('off_route_early', 3.0), ('off_route', 9.0), ('off_route_again', 10.0),
            ('few_depth', 16.0), ('off_route_second_site', 22.0), ('back_at_new_site', 30.0), ('back_at_new_site', 31.0)]
    for i, (kind, ts) in enumerate(plan):
        fr = synth.make_frame(scene, 5600 + i, k_star=0, n_desc=500, n_planted=260)
        desc, pts = fr['desc'], fr['pts2d']
        if kind == 'near_route':
            k = 7
            fr = synth.make_frame(data, 5650, k_star=k, n_desc=500, n_planted=200)
            desc, pts = fr['desc'], fr['pts2d']
            lm = data['landmarks'][k]
            bx, by, yaw = lm['pose'][0] - 0.35, lm['pose'][1], 0.0
        elif kind == 'off_route_second_site':
            bx, by, yaw = 10.0, -42.0, 0.4
        else:
            bx, by, yaw = 12.0 + 0.05 * i, 30.0, 0.1
        # depth image (mm): the planted points' depths at their pixels, a wall elsewhere; some zeros
        depth = np.full((480, 640), 4000, dtype=np.uint16)
        if kind == 'few_depth':
            depth[:] = 200                                  # closer than 0.5 m: fewer than ACCUM_MIN_KPTS valid points
        uu = np.clip(np.round(pts[:, 0]).astype(int), 0, 639)
        vv = np.clip(np.round(pts[:, 1]).astype(int), 0, 479)
        if kind != 'few_depth':
            depth[vv, uu] = rng.integers(400, 16500, len(uu)).astype(np.uint16)
            if kind != 'near_route':
                # planted keypoints: the depth the scene point really has in THIS camera, so that the landmark
                # accumulated from this frame is geometrically consistent with the later frames of the same place
                X = scene['landmarks'][0]['keypoints_3d_cam'][fr['t_idx']].astype(np.float64)
                Xc = X @ synth.rodrigues(fr['rvec']).T + np.asarray(fr['tvec']).reshape(1, 3)
                z_mm = np.clip(np.round(Xc[:, 2] * 1000.0), 0, 65535).astype(np.uint16)
                depth[vv[fr['q_slots']], uu[fr['q_slots']]] = z_mm
            depth[vv[::17], uu[::17]] = 0
        node.last_depth = depth
        base_pose = (float(bx), float(by), 0.0, 0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2)))
        node.orb.next = (desc, pts)
        node._read_pose = (lambda bp=base_pose: bp)
        real_time = vm.time.time
        vm.time.time = (lambda t=ts: 1000.0 + t)            # the node stamps ticks with time.time()
        n_lm_before = len(node.landmarks)
        n_before = len(node.anchor_pub.sent)
        try:
            node._tick()
        finally:
            vm.time.time = real_time
        line = open(csv).read().strip().split('\n')[-1].split(',')
        rec = {'kind': kind, 'ts': 1000.0 + ts, 'base_pose': base_pose, 'csv': line[1:], 'depth': depth,
               'published': len(node.anchor_pub.sent) > n_before, 'n_landmarks': len(node.landmarks),
               'appended': len(node.landmarks) > n_lm_before}
        if rec['published']:
            m = node.anchor_pub.sent[-1]
            p_, o_ = m.pose.pose.position, m.pose.pose.orientation
            rec['anchor'] = [p_.x, p_.y, p_.z, o_.x, o_.y, o_.z, o_.w]
        ticks.append((rec, desc, pts))
    # the SIGTERM handler the node registered: writes <pkl>_augmented.pkl and exits
    try:
        signal.getsignal(signal.SIGTERM)()
    except SystemExit:
        pass
    aug = pickle.load(open(pkl.replace('.pkl', '_augmented.pkl'), 'rb'))
    new = [lm for lm in aug['landmarks'] if lm.get('accumulated')]
    assert len(new) == node.n_accumulated and len(new) >= 2, (len(new), node.n_accumulated)
    np.savez_compressed(
        os.path.join(out_dir, 'accum_golden.npz'), lib_seed=np.int64(56),
        kinds=np.array([t[0]['kind'] for t in ticks]), ts=np.array([t[0]['ts'] for t in ticks]),
        base_pose=np.array([t[0]['base_pose'] for t in ticks]),
        csv=np.array([','.join(t[0]['csv']) for t in ticks]), published=np.array([t[0]['published'] for t in ticks]),
        appended=np.array([t[0]['appended'] for t in ticks]), n_landmarks=np.array([t[0]['n_landmarks'] for t in ticks]),
        anchor=np.array([t[0].get('anchor', [0] * 7) for t in ticks], dtype=np.float64),
        desc=np.stack([t[1] for t in ticks]), pts2d=np.stack([t[2] for t in ticks]),
        depth=np.stack([t[0]['depth'] for t in ticks]),
        n_aug_landmarks=np.int64(len(aug['landmarks'])), aug_keys=np.array(sorted(aug.keys())),
        new_pose=np.array([lm['pose'] for lm in new], dtype=np.float64), new_ts=np.array([lm['ts'] for lm in new]),
        new_n=np.array([lm['n_features'] for lm in new], dtype=np.int64),
        new_desc=np.concatenate([lm['descriptors'] for lm in new]),
        new_kp2d=np.concatenate([lm['keypoints_2d'] for lm in new]),
        new_kp3d=np.concatenate([lm['keypoints_3d_cam'] for lm in new]))
    print('accum_golden.npz', [t[0]['csv'][-1] for t in ticks], 'landmarks', [t[0]['n_landmarks'] for t in ticks],
          'new n_features', [lm['n_features'] for lm in new])


def _lift_depth(rng, kind):
    """Synthetic aligned depth in millimetres: a ground ramp below the horizon, boxes at other ranges
    (depth discontinuities), holes (0), far background beyond the 15 m gate."""
    v = np.arange(480, dtype=np.float64)[:, None]
    u = np.arange(640, dtype=np.float64)[None, :]
    z = np.where(v > 245, 0.48 * 320.0 / np.maximum(v - 240.0, 1.0), 40.0) + 0.0 * u        # ground plane seen from 0.48 m
    for _ in range(12):
        u0, v0 = int(rng.integers(0, 600)), int(rng.integers(150, 440))
        w, h = int(rng.integers(20, 120)), int(rng.integers(20, 120))
        z[v0:v0 + h, u0:u0 + w] = rng.uniform(0.3, 18.0)
    z = z + rng.normal(0, 0.004, z.shape)
    mm = np.clip(z * 1000.0, 0, 65535).astype(np.uint16)
    mm[rng.random(mm.shape) < (0.25 if kind == 'holes' else 0.03)] = 0
    if kind == 'far':
        mm[:] = 30000
    return mm


def golden_lift(out_dir):
    """The teach-time recorder: VisualLandmarkRecorder._tick (visual_landmark_recorder.py:211-306) run unmodified
    under ROS stubs with a stand-in ORB; every tick's inputs and the `landmarks` record it appended (or not)."""
    import importlib
    import pickle
    ros_stubs.import_reference()                      # installs the stubs and the scripts/common path
    import sys
    saved = sys.argv
    sys.argv = ['x']
    try:
        rec_mod = importlib.import_module('visual_landmark_recorder')
    finally:
        sys.argv = saved
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'out', 'landmarks.pkl')
    node = rec_mod.VisualLandmarkRecorder(pkl, min_disp_m=2.0)

    class _Kp:
        def __init__(self, pt):
            self.pt = (float(pt[0]), float(pt[1]))

    class _Orb:
        def detectAndCompute(self, gray, mask):
            d, p = self.next
            return [_Kp(q) for q in p], d

    node.orb = _Orb()
    node.last_rgb = np.zeros((480, 640, 3), dtype=np.uint8)
    rng = np.random.default_rng(2026)
    kinds = ['ok', 'near', 'ok', 'holes', 'far', 'ok', 'few', 'ok']
    x = 0.0
    ticks = []
    for i, kind in enumerate(kinds):
        x += 0.5 if kind == 'near' else 3.0          # 'near': displacement < min_disp -> tick returns early
        n = 40 if kind == 'few' else 500
        pts = np.stack([rng.uniform(-3, 643, n), rng.uniform(100, 483, n)], axis=1).astype(np.float32)
        pts[:25, 0] = np.round(pts[:25, 0]) + 0.5     # exact halves: np.round is half-to-even
        pts[25:50, 1] = np.round(pts[25:50, 1]) + 0.5
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        depth = _lift_depth(rng, kind)
        node.orb.next = (desc, pts)
        node.last_depth = depth
        node.last_rgb_ts = 100.0 + i
        base = (x, 0.1 * i, 0.0, 0.0, 0.0, float(np.sin(0.05 * i)), float(np.cos(0.05 * i)))
        node._read_pose = (lambda bp=base: bp)
        before = len(node.landmarks)
        node._tick()
        rec = node.landmarks[-1] if len(node.landmarks) > before else None
        ticks.append((kind, pts, desc, depth, base, rec))
    node._save()
    saved_pkl = pickle.load(open(pkl, 'rb'))
    assert len(saved_pkl['landmarks']) == len(node.landmarks)
    NMAX = 500
    def pad(a, shape, dtype):
        out = np.zeros(shape, dtype=dtype)
        out[:len(a)] = a
        return out
    np.savez_compressed(
        os.path.join(out_dir, 'lift_golden.npz'),
        kinds=np.array([t[0] for t in ticks]),
        n_kpts=np.array([len(t[1]) for t in ticks], dtype=np.int32),
        kpts=np.stack([pad(t[1], (NMAX, 2), np.float32) for t in ticks]),
        desc=np.stack([pad(t[2], (NMAX, 32), np.uint8) for t in ticks]),
        depth=np.stack([t[3] for t in ticks]),
        base_pose=np.array([t[4] for t in ticks], dtype=np.float64),
        recorded=np.array([t[5] is not None for t in ticks]),
        n_feat=np.array([t[5]['n_features'] if t[5] else 0 for t in ticks], dtype=np.int32),
        cam_pose=np.array([t[5]['pose'] if t[5] else [0] * 7 for t in ticks], dtype=np.float64),
        ts=np.array([t[5]['ts'] if t[5] else 0.0 for t in ticks], dtype=np.float64),
        rec_desc=np.stack([pad(t[5]['descriptors'], (NMAX, 32), np.uint8) if t[5] else np.zeros((NMAX, 32), np.uint8) for t in ticks]),
        rec_kp2d=np.stack([pad(t[5]['keypoints_2d'], (NMAX, 2), np.float32) if t[5] else np.zeros((NMAX, 2), np.float32) for t in ticks]),
        rec_kp3d=np.stack([pad(t[5]['keypoints_3d_cam'], (NMAX, 3), np.float32) if t[5] else np.zeros((NMAX, 3), np.float32) for t in ticks]),
        pkl_keys=np.array(sorted(saved_pkl.keys())),
        pkl_intrinsics=np.array([saved_pkl['intrinsics'][k] for k in ('fx', 'fy', 'cx', 'cy', 'width', 'height')], dtype=np.float64),
        pkl_b2c_t=np.array(saved_pkl['base_to_cam_translation'], dtype=np.float64),
        pkl_b2c_R=np.array(saved_pkl['base_to_cam_rot'], dtype=np.float64))
    print('lift_golden.npz', [(t[0], t[5]['n_features'] if t[5] else None) for t in ticks])


def golden_reloc(out_dir):
    """exp 63 global relocalisation: experiments/63_global_reloc/scripts/visual_landmark_matcher.py::_tick run
    unmodified under ROS stubs (stand-in ORB; '/tmp/drift_est.txt' served from memory).  Ticks far from every
    landmark with a large drift estimate trigger the whole-library search."""
    import builtins
    import importlib.util
    import pickle
    ros_stubs.import_reference()
    import sys
    path = os.path.join(ros_stubs.REFERENCE_COMMON, '..', '..', 'experiments', '63_global_reloc', 'scripts',
                        'visual_landmark_matcher.py')
    saved = sys.argv
    sys.argv = ['x']
    try:
        spec = importlib.util.spec_from_file_location('visual_landmark_matcher_exp63', os.path.normpath(path))
        vm = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(vm)
    finally:
        sys.argv = saved
    vm.ACCUM_ENABLE = False
    drift = {'value': 0.0}

    def fake_open(name, *a, **k):
        if name == '/tmp/drift_est.txt':
            return io.StringIO(f"{drift['value']}\n")
        return builtins.open(name, *a, **k)

    vm.open = fake_open
    import types
    clock = {'ts': 0.0}
    vm.time = types.SimpleNamespace(time=lambda: clock['ts'])
    data = synth.make_library(63, n_kf=60, n_desc=300, ragged=True, route_len_m=120.0)
    tmp = tempfile.mkdtemp()
    pkl = os.path.join(tmp, 'south_landmarks.pkl')
    with open(pkl, 'wb') as f:
        pickle.dump(data, f)
    csv = os.path.join(tmp, 'log', 'anchor_matches.csv')
    node = vm.VisualLandmarkMatcher(pkl, csv)
    node.last_rgb = np.zeros((480, 640, 3), dtype=np.uint8)
    node.last_depth = np.full((480, 640), 2000, dtype=np.uint16)

    class _Kp:
        def __init__(self, pt):
            self.pt = (float(pt[0]), float(pt[1]))

    class _Orb:
        def detectAndCompute(self, gray, mask):
            d, p = self.next
            return [_Kp(q) for q in p], d

    node.orb = _Orb()
    rng = np.random.default_rng(63)
    # (kind, drift estimate, ts): 'lost' = robot reported 60 m off the route, frame really taken at landmark k
    plan = [('local', 0.0, 30.0), ('lost_nodrift', 0.5, 60.0), ('lost', 8.0, 90.0), ('lost_recent', 8.0, 95.0),
            ('lost', 6.0, 130.0), ('lost_random', 9.0, 170.0), ('lost', 12.0, 210.0), ('lost_weak', 7.0, 250.0)]
    ticks = []
    for i, (kind, dr, ts) in enumerate(plan):
        k = int(rng.integers(5, 55))
        planted = 0 if kind == 'lost_random' else (60 if kind == 'lost_weak' else 200)
        fr = synth.make_frame(data, 6300 + i, k_star=k, n_desc=500, n_planted=planted)
        lm = data['landmarks'][k]
        bx, by = lm['pose'][0] - 0.35 + rng.normal(0, 0.3), lm['pose'][1] + rng.normal(0, 0.3)
        if kind != 'local':
            by += 60.0
        yaw = rng.normal(0, 0.1)
        base_pose = (float(bx), float(by), 0.0, 0.0, 0.0, float(np.sin(yaw / 2)), float(np.cos(yaw / 2)))
        node.orb.next = (fr['desc'], fr['pts2d'])
        node._read_pose = (lambda bp=base_pose: bp)
        drift['value'] = dr
        clock['ts'] = ts                              # the node stamps ticks with time.time() (line 291)
        n_before = len(node.anchor_pub.sent)
        node._tick()
        line = open(csv).read().strip().split('\n')[-1].split(',')
        rec = {'kind': kind, 'base_pose': base_pose, 'csv': line, 'published': len(node.anchor_pub.sent) > n_before,
               'drift': dr, 'k': k}
        if rec['published']:
            m = node.anchor_pub.sent[-1]
            p, o = m.pose.pose.position, m.pose.pose.orientation
            rec['anchor'] = [p.x, p.y, p.z, o.x, o.y, o.z, o.w]
        ticks.append((rec, fr['desc'], fr['pts2d']))
    np.savez_compressed(
        os.path.join(out_dir, 'reloc_golden.npz'), lib_seed=np.int64(63),
        kinds=np.array([t[0]['kind'] for t in ticks]), base_pose=np.array([t[0]['base_pose'] for t in ticks]),
        drift=np.array([t[0]['drift'] for t in ticks]), k_true=np.array([t[0]['k'] for t in ticks], dtype=np.int32),
        ts=np.array([float(t[0]['csv'][0]) for t in ticks]),
        csv=np.array([','.join(t[0]['csv'][1:]) for t in ticks]), published=np.array([t[0]['published'] for t in ticks]),
        anchor=np.array([t[0].get('anchor', [0] * 7) for t in ticks], dtype=np.float64),
        desc=np.stack([t[1] for t in ticks]), pts2d=np.stack([t[2] for t in ticks]))
    print('reloc_golden.npz', [(t[0]['kind'], t[0]['csv'][0], t[0]['csv'][-1]) for t in ticks])


def golden_hitcount(out_dir):
    """datasets/rover/scripts/occupancy_astar.py::build_occupancy (and backproject_depth for realistic inputs), imported
    unmodified with matplotlib stubbed out (it is only used for the figures)."""
    import importlib.util
    import sys
    import types
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches'):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.use = lambda *a, **k: None
            sys.modules[name] = m
    path = os.path.normpath(os.path.join(ros_stubs.REFERENCE_COMMON, '..', '..', '..', '..', 'datasets', 'rover', 'scripts',
                                         'occupancy_astar.py'))
    # The script as shipped does not import: line 46 (`y_rel = y_point - y_camera`) is a comment that lost its '#'.
    # Its top-level statements are therefore executed one by one and the ones that raise NameError are skipped
    # (exactly that line); every function and constant is the reference's own, untouched.
    import ast
    src = open(path).read()
    ns = {'__name__': 'occupancy_astar_ref', '__file__': path}
    skipped = []
    for node in ast.parse(src, path).body:
        try:
            exec(compile(ast.Module([node], []), path, 'exec'), ns)
        except NameError as e:
            skipped.append((node.lineno, str(e)))
    assert [ln for ln, _ in skipped] == [46], skipped
    oa = types.SimpleNamespace(**ns)
    ns['log'] = lambda *a, **k: None            # the functions look `log` up in their globals
    rng = np.random.default_rng(44)
    pts_all, lab_all = [], []
    for f in range(12):                                  # a short walk through a synthetic garden, RealSense intrinsics
        v = np.arange(480, dtype=np.float64)[:, None]
        z = np.where(v > 260, 0.6 * 593.14 / np.maximum(v - 245.16, 1.0), 9.0) + np.zeros((480, 640))
        for _ in range(6):
            u0, v0 = int(rng.integers(0, 600)), int(rng.integers(60, 400))
            z[v0:v0 + int(rng.integers(30, 150)), u0:u0 + int(rng.integers(20, 90))] = rng.uniform(0.4, 5.5)
        depth = np.clip((z + rng.normal(0, 0.01, z.shape)) * 1000.0, 0, 65535).astype(np.uint16)
        depth[rng.random(depth.shape) < 0.05] = 0
        T = np.eye(4)
        yaw = 0.15 * f
        T[:3, :3] = np.array([[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]])
        T[:3, 3] = [0.4 * f, -0.6, 0.3 * f]
        p, l = oa.backproject_depth(depth, T, oa.PIXEL_STEP)
        pts_all.append(p)
        lab_all.append(l)
    points = np.concatenate(pts_all)
    labels = np.concatenate(lab_all)
    occ, x_min, z_min, nx, nz = oa.build_occupancy(points, labels, oa.GRID_RES)
    np.savez_compressed(os.path.join(out_dir, 'hitcount_golden.npz'), points=points, labels=labels,
                        occupancy=occ, x_min=np.float64(x_min), z_min=np.float64(z_min), nx=np.int64(nx), nz=np.int64(nz),
                        grid_res=np.float64(oa.GRID_RES), min_total=np.int64(oa.MIN_HITS_TOTAL),
                        min_obstacle=np.int64(oa.MIN_HITS_OBSTACLE))
    print('hitcount_golden.npz', points.shape, (nz, nx), {int(k): int((occ == k).sum()) for k in (-1, 0, 1)})
