"""Drop-in for the reference's eval_ycb.py: scoring of tracked YCB-Video sequences.

  VOCap            eval_ycb.py:45-64    area under the accuracy-vs-threshold curve below 0.1 m, on the GPU (sort + one reduction)
  eval_one_class   eval_ycb.py:67-119   every pose file under --res_dir (what predict.getResultsYcb writes: seq<id>/%07d.txt) that falls
                                        on a keyframe, against <ycb_dir>/data_organized/%04d/pose_gt/<class>/%06d.txt, with the class's
                                        CADmodels/*/points.xyz: ADD and ADD-S of ALL key frames in one launch (the reference loops
                                        Utils.add / Utils.adi per frame), then their AUCs
  eval_all         eval_ycb.py:121-162  the 21 classes' result folders under one root, pooled

Same file conventions, same printed lines, same return values; `python -m <package>.eval_ycb --ycb_dir .. --res_dir .. --class_id ..`.
"""
import argparse, glob, os
import numpy as np
import torch
from . import Utils as U


def VOCap(rec):
    eng = U._eng()
    errs = torch.from_numpy(np.ascontiguousarray(np.asarray(rec, dtype=np.float64).reshape(-1))).to(eng.device)
    return eng.vocap(errs)


def _read_points(path):
    with open(path, 'r') as ff:
        return np.array([list(map(float, line.rstrip().split())) for line in ff if line.strip()], dtype=np.float64).reshape(-1, 3)


def eval_one_class(args):
    pose_files = sorted(glob.glob(args.res_dir + '**/*.txt', recursive=True))
    assert len(pose_files) > 0, 'args.res_dir is\n{}'.format(args.res_dir)
    class_names = sorted(os.listdir('{}/CADmodels/'.format(args.ycb_dir)))
    model_files = sorted(glob.glob('{}/CADmodels/**/points.xyz'.format(args.ycb_dir), recursive=True))
    model_pts = _read_points(model_files[args.class_id - 1])
    with open('{}/YCB_Video_toolbox/keyframe.txt'.format(args.ycb_dir), 'r') as ff:
        keyframes = set(line.rstrip() for line in ff.readlines())

    preds, gts = [], []
    for pose_file in pose_files:
        seq_id = int(pose_file.replace(args.res_dir, '').split('/')[0].replace('seq', ''))
        frame_id = int(os.path.basename(pose_file).split('.')[0]) + 1
        if '%04d/%06d' % (seq_id, frame_id) not in keyframes:
            continue
        preds.append(np.loadtxt(pose_file))
        gts.append(np.loadtxt('{}/data_organized/%04d/pose_gt/{}/%06d.txt'.format(args.ycb_dir, args.class_id) % (seq_id, frame_id)))
    assert len(preds) > 0
    eng = U._eng()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(eng.device)
    add_d, adi_d = eng.add_adi(t(model_pts), t(np.stack(preds).reshape(-1, 4, 4)), t(np.stack(gts).reshape(-1, 4, 4)))
    adi_errs = np.sort(adi_d.cpu().numpy())
    add_errs = np.sort(add_d.cpu().numpy())

    add_aps = VOCap(add_errs) * 100
    print('>>>>>>>>>>>>>>>> args.class_id:', args.class_id, class_names[args.class_id - 1])
    print('add:', add_aps)
    adi_aps = VOCap(adi_errs) * 100
    print('adi:', adi_aps)
    return adi_errs, add_errs


def eval_all(args):
    class_ids = np.arange(1, 22)
    print(class_ids)
    root = getattr(args, 'res_root', None) or '/home/bowen/debug/Ours/'           # the reference hard-codes this path (eval_ycb.py:125)
    if not root.endswith('/'):
        root += '/'
    res_dirs = []
    for class_folder in sorted(os.listdir(root)):
        for folder in os.listdir(root + class_folder):
            if os.path.isdir(root + class_folder + '/' + folder):
                res_dirs.append(root + class_folder + '/' + folder + '/')
                break
    for res_dir in res_dirs:
        print(res_dir)
    assert len(res_dirs) == len(class_ids), 'len(res_dirs)={}'.format(len(res_dirs))
    adi_errs, add_errs = [], []
    for i, class_id in enumerate(class_ids):
        args.res_dir = res_dirs[i]
        args.class_id = int(class_id)
        res = eval_one_class(args)
        adi_errs += list(res[0])
        add_errs += list(res[1])
    adi_errs, add_errs = np.array(adi_errs), np.array(add_errs)
    n = len(adi_errs)
    expected = getattr(args, 'expected_total', None)                              # 14025 key-frame poses on the real test set (eval_ycb.py:153)
    assert expected is None or n == expected
    add_aps = VOCap(add_errs) * 100
    print()
    print('add:', add_aps)
    adi_aps = VOCap(adi_errs) * 100
    print('adi:', adi_aps)
    print('Total res num:', n)
    return adi_aps, add_aps, n


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--ycb_dir', required=True)
    parser.add_argument('--class_id', type=int, default=1)
    parser.add_argument('--res_dir', type=str, default=None, help='one class: the folder with seq<id>/%%07d.txt')
    parser.add_argument('--res_root', type=str, default=None, help='all classes: <res_root>/<class folder>/<run folder>/seq<id>/%%07d.txt')
    parser.add_argument('--expected_total', type=int, default=None)
    args = parser.parse_args(argv)
    if args.res_root is not None:
        return eval_all(args)
    if args.res_dir is None:
        parser.error('need --res_dir (one class) or --res_root (all classes)')
    return eval_one_class(args)


if __name__ == '__main__':
    main()
