#!/usr/bin/env python
"""Job driver with the flags of the reference's ``tools/occ/occ_annotate.py`` (:203-225, 649-671) on top of the
batched CUDA path: annotate every tracklet of a converted Waymo directory and write
``<out-dir>/<split>/<segment>/<id>.npz``.

    python tools/occ_annotate_job.py --data-root data/waymo --out-dir out --split training \\
        --tracklets vehicle_tracklets.npz --voxel-size 0.2 [--object-type vehicle] [--overwrite] [--save-mean-var]
    python -m torch.distributed.run --nproc-per-node 8 tools/occ_annotate_job.py ...      # one process per GPU

Differences from the reference CLI: tracklets come from a plain npz file (``waymo_io.save_tracklet_records``)
instead of a Waymo ``.bin`` + pickled ``LiDARTracklet`` cache (both need packages outside this repo);
``--workers`` / ``--ngpus`` are replaced by the launcher's world size (segments are dealt to ranks the way the
reference deals them to workers); ``--cpu-voxelization`` and ``--debug`` select dead branches in the reference
and do not exist here.
"""
import argparse
import functools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

TYPE_MAPPING = {"vehicle": 1, "pedestrian": 2, "cyclist": 3}          # occ_annotate.py:229-233


def main(argv=None, annotate_fn=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-root", required=True)
    ap.add_argument("--out-dir", required=True)
    ap.add_argument("--split", default="training")
    ap.add_argument("--voxel-size", type=float, default=0.2)
    ap.add_argument("--tracklets", required=True, help="npz written by waymo_io.save_tracklet_records")
    ap.add_argument("--object-type", default="vehicle", choices=sorted(TYPE_MAPPING))
    ap.add_argument("--overwrite", action="store_true")
    ap.add_argument("--save-mean-var", action="store_true")
    args = ap.parse_args(argv)

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if annotate_fn is None:
        import torch

        from objectcentricocccompletion_b200 import occ_annotate

        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        annotate_fn = functools.partial(occ_annotate.annotate_batch, save_mean_var=args.save_mean_var)
    from objectcentricocccompletion_b200 import waymo_io

    records = [r for r in waymo_io.load_tracklet_records(args.tracklets) if r.type == TYPE_MAPPING[args.object_type]]
    mine = set(waymo_io.segments_of_rank([r.segment_name for r in records], rank, world))
    records = [r for r in records if r.segment_name in mine]
    paths = waymo_io.annotate_from_disk(records, args.data_root, args.out_dir, args.split, args.voxel_size,
                                        args.overwrite, annotate_fn)
    n = sum(p is not None for p in paths)
    print(f"[rank {rank}/{world}] {len(mine)} segments, {len(records)} tracklets, {n} files under {args.out_dir}")
    return paths


if __name__ == "__main__":
    main()
