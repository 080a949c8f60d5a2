import sys; sys.path.insert(0,'.')
import numpy as np, torch
import objectcentricocccompletion_b200 as occ
from objectcentricocccompletion_b200 import synth
from oracle import oracle
from oracle.make_golden import edge_batch
b = edge_batch()
got = occ.annotate_batch(b); exp = oracle.annotate_batch(b)
assert all((g['occ']==e['occ']).all() for g,e in zip(got,exp) if e['occ'] is not None)
b = synth.make_batch(2, 10, 0.2, seed=9, small=True)
for f in (0,1,2): occ.annotate_batch(b, flags=f)
pts, bidx = synth.scatter_inputs(4, 4, 256, 5, seed=1)
p = torch.from_numpy(pts).cuda()
vs, pcr = [0.2]*3, [-204.8,-204.8,-4,204.8,204.8,8]
c = occ.Voxelization(vs, pcr, -1)(p)
f = p[:, :3].contiguous().requires_grad_()
o, oc = occ.DynamicScatter(vs, pcr, False)(f, c); o.sum().backward()
c4 = torch.cat([torch.from_numpy(bidx).cuda()[:,None].int(), c],1).contiguous()
occ.DynamicScatter(vs, pcr, True)(p[:, :3].contiguous(), c4)
occ.scatter_v2(p, c4.long(), 'max')
occ.voxelization(p, [0.5]*3, [-210,-210,-5,210,210,9], 8, 3000)
occ.points_in_boxes_gpu(p[None,:,:3].contiguous(), torch.tensor([[[0,0,0,4,4,4,0.3]]],device='cuda'))
torch.cuda.synchronize(); print('sanitizer workload ok')
# round-1 additions: save-mean-var, range-image builder, label mirror, tracklet point extraction, graph replay
from objectcentricocccompletion_b200 import occ_annotate, range_image, track_input, occ_ops
b = synth.make_batch(3, 12, 0.2, seed=1, small=True)
r = occ.annotate_batch(b, save_mean_var=True)
assert all(x["mean_var"] is not None for x in r if x["occ"] is not None)
seg = b.segments[0]
pts = [np.concatenate([t.points[0][:, :3] for t in b.tracklets], 0).astype(np.float32)] * 2
range_image.build_range_images(pts, seg.extrinsics[0, :2], [seg.inclinations[0], seg.inclinations[1]],
                               [seg.range_images[0][0].shape, seg.range_images[1][0].shape])
occ_ops.mirror_occ_label([torch.from_numpy(x["occ"]).cuda() for x in r if x["occ"] is not None])
track_input.crop_frame(pts[0], b.tracklets[0].boxes[:3], 0.5)
pk = occ_annotate.pack_tracklets(b)
d = occ_annotate.DeviceTracklets(pk); d.upload(occ_annotate.HostBuffers(pk)); d.capture(); d.replay()
torch.cuda.synchronize(); print('sanitizer workload (round-1 additions) ok')
# round-2 additions: the three upload modes (window marks, pull, tile flags, host gather + scatter), brick-cull
# on / off, point pool, candidate selection from whole-frame clouds
from objectcentricocccompletion_b200 import candidates, point_pool
b = synth.make_batch(3, 12, 0.2, seed=2, small=True)
pk = occ_annotate.pack_tracklets(b)
for pin, win in ((True, True), (False, True), (True, False)):
    d = occ_annotate.DeviceTracklets(pk, labels="both"); d.upload(occ_annotate.HostBuffers(pk, pin=pin, windows=win))
    for fl in (0, occ_annotate.FLAG_NO_BRICK_CULL, occ_annotate.FLAG_TINY_QUEUE, occ_annotate.FLAG_CUDA_ARITH):
        d.run(fl)
    torch.cuda.synchronize()
rois = torch.tensor([[0, 0, 0, 4, 2, 2, 0.3], [5, 5, 0, 4, 2, 2, 1.0]], device='cuda')
pp = torch.randn(2000, 3, device='cuda') * 3
point_pool.dynamic_point_pool_mixed(rois, torch.tensor([0, 1], device='cuda'), pp, torch.randint(0, 2, (2000,), device='cuda'),
                                    [0.5, 0.5, 0.5], 64)
clouds = [np.random.default_rng(0).normal(size=(5000, 3)).astype(np.float32) * 10 for _ in range(3)]
candidates.select_candidates(clouds, [np.array([[0, 0, 0, 4, 2, 2, 0.1]] * 3, np.float32)], [np.arange(3)], 1.0)
torch.cuda.synchronize(); print('sanitizer workload (round-2 additions) ok')
