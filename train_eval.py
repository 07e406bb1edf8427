#!/usr/bin/env python
"""Re-train a searched TF-NAS architecture — same command line as the reference ``train_eval.py`` (flags :29-58, epoch loop
:197-240) plus the mixed-precision / distributed switches of ``train_eval_amp.py`` in the form this image supports.

    python train_eval.py --model_path searched_model_90.pth.tar --train_root ... --val_root ... --train_list ... --val_list ...
    python train_eval.py --config_path model.config --synthetic 50 --amp bf16 --channels_last
    torchrun --nproc-per-node 8 train_eval.py ...          # one process per GPU, NCCL gradient all-reduce

The network comes from ``tfnas_b200.model_eval`` (``Network`` from a search checkpoint, ``NetworkCfg`` from a
``model.config`` JSON); the per-step glue (label-smoothing cross-entropy, clip + SGD) runs in the library's fused kernels.
``--batch_size`` is the GLOBAL batch as in ``train_eval_amp.py:193`` (each rank takes batch_size / world)."""
import argparse
import json
import logging
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tfnas_b200 import config as cfg  # noqa: E402
from tfnas_b200 import eval_loop, model_eval, parallel, parsing  # noqa: E402
from tfnas_b200.lut import load_lut  # noqa: E402
from train_search import SyntheticQueue, cosine_lr_list  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser('training the searched architecture on imagenet')
    p.add_argument('--train_root', type=str, default=None, help='training image root path')
    p.add_argument('--val_root', type=str, default=None, help='validating image root path')
    p.add_argument('--train_list', type=str, default=None, help='training image list')
    p.add_argument('--val_list', type=str, default=None, help='validating image list')
    p.add_argument('--model_path', type=str, default='', help='the searched model path')
    p.add_argument('--config_path', type=str, default='', help='the model config path')
    p.add_argument('--save', type=str, default='./checkpoints/', help='model and log saving path')
    p.add_argument('--snapshot', type=str, default='', help='for reset')
    p.add_argument('--print_freq', type=float, default=100)
    p.add_argument('--workers', type=int, default=16)
    p.add_argument('--epochs', type=int, default=250)
    p.add_argument('--batch_size', type=int, default=512)
    p.add_argument('--lr', type=float, default=0.2)
    p.add_argument('--momentum', type=float, default=0.9)
    p.add_argument('--weight_decay', type=float, default=1e-5)
    p.add_argument('--grad_clip', type=float, default=5.0)
    p.add_argument('--label_smooth', type=float, default=0.1)
    p.add_argument('--num_classes', type=int, default=1000)
    p.add_argument('--dropout_rate', type=float, default=0.2)
    p.add_argument('--drop_connect_rate', type=float, default=0.2)
    p.add_argument('--seed', type=int, default=2)
    p.add_argument('--note', type=str, default='try')
    # additions
    p.add_argument('--amp', type=str, default='none', choices=['none', 'bf16'], help='bf16 autocast (train_eval_amp.py role)')
    p.add_argument('--channels_last', action='store_true')
    p.add_argument('--synthetic', type=int, default=0, help='use N synthetic batches per epoch (no dataset needed)')
    p.add_argument('--image_size', type=int, default=224)
    p.add_argument('--lookup_path', type=str, default='', help='latency table: log the derived network\'s table latency')
    return p


def build_model(args, lat_lookup=None):
    """train_eval.py:103-115: a search checkpoint is parsed into an architecture; a config file is instantiated as is."""
    if args.model_path and os.path.isfile(args.model_path):
        op_w, depth_w = parsing.get_op_and_depth_weights(args.model_path)
        parsed_arch = parsing.parse_architecture(op_w, depth_w)
        masks = torch.load(args.model_path, map_location='cpu', weights_only=False)['mc_mask_dddict']
        return model_eval.Network(args.num_classes, parsed_arch, cfg.get_mc_num_dddict(masks), lat_lookup,
                                  args.dropout_rate, args.drop_connect_rate)
    if args.config_path and os.path.isfile(args.config_path):
        with open(args.config_path) as f:
            return model_eval.NetworkCfg(args.num_classes, json.load(f), lat_lookup, args.dropout_rate,
                                         args.drop_connect_rate)
    raise SystemExit('invalid --model_path and --config_path')


def make_queues(args, rank, world):
    bs = args.batch_size // world
    if args.synthetic > 0:
        return (SyntheticQueue(args.synthetic, bs, args.image_size, args.num_classes, args.seed + rank),
                SyntheticQueue(max(1, args.synthetic // 2), bs, args.image_size, args.num_classes, 1000 + args.seed + rank))
    if not (args.train_root and args.val_root and args.train_list and args.val_list):
        raise SystemExit('--train_root/--val_root/--train_list/--val_list are required unless --synthetic N is given')
    from tfnas_b200.data import make_eval_loaders
    return make_eval_loaders(args, rank, world)


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, local, world = parallel.init_from_env()
    if not torch.cuda.is_available():
        logging.info('No GPU device available')
        sys.exit(1)
    torch.cuda.set_device(local)
    args.save = os.path.join(args.save, 'eval-{}-{}'.format(time.strftime('%Y%m%d-%H%M%S'), args.note))
    if rank == 0:
        os.makedirs(args.save, exist_ok=True)
    fmt = '%(asctime)s %(message)s'
    logging.basicConfig(stream=sys.stdout, level=logging.INFO if rank == 0 else logging.WARNING, format=fmt,
                        datefmt='%m/%d %I:%M:%S %p')
    logging.getLogger().setLevel(logging.INFO if rank == 0 else logging.WARNING)
    if rank == 0:
        fh = logging.FileHandler(os.path.join(args.save, 'log.txt'))
        fh.setFormatter(logging.Formatter(fmt))
        logging.getLogger().addHandler(fh)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    logging.info('args = %s', args)
    bench_flag = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True                    # train_eval.py:99; restored for in-process callers (tests)
    try:
        _run(args, rank, world)
    finally:
        torch.backends.cudnn.benchmark = bench_flag
        if rank == 0:
            logging.getLogger().removeHandler(fh)
            fh.close()


def _run(args, rank, world):

    logging.info('parsing the architecture')
    lat_lookup = load_lut(args.lookup_path) if args.lookup_path else None
    model = build_model(args, lat_lookup).cuda()
    eval_loop.broadcast_model(model)
    if rank == 0:
        with open(os.path.join(args.save, 'model.config'), 'w') as f:
            json.dump(model.config, f, indent=4)
    logging.info('param size = %fMB', sum(np.prod(v.size()) for v in model.parameters()) / 1e6)
    if lat_lookup:
        logging.info('table latency = %f', model.get_lookup_latency(224))      # the table is indexed by 224x224 feature-map sizes

    criterion_smooth, criterion = eval_loop.make_criteria(args.label_smooth)
    optimizer = eval_loop.make_optimizer(model, args.lr, args.momentum, args.weight_decay)
    train_queue, val_queue = make_queues(args, rank, world)
    sync = parallel.GradSync()
    lr_list = cosine_lr_list(args.lr, args.epochs)
    best1 = best5 = 0.0
    start_epoch = 0
    if args.snapshot and os.path.isfile(args.snapshot):
        logging.info('loading snapshot from {}'.format(args.snapshot))
        ck = torch.load(args.snapshot, map_location='cuda', weights_only=False)
        start_epoch, best1, best5 = ck['epoch'], ck['best_acc_top1'], ck['best_acc_top5']
        model.load_state_dict({k[7:] if k.startswith('module.') else k: v for k, v in ck['state_dict'].items()})
        optimizer.load_state_dict(ck['optimizer'])

    for epoch in range(start_epoch, args.epochs):
        lr = eval_loop.epoch_lr(lr_list, epoch, args.batch_size)
        logging.info('Epoch: %d lr %e', epoch, lr_list[epoch])
        if lr != lr_list[epoch]:
            logging.info('Warming-up Epoch: %d, LR: %e', epoch, lr)
        eval_loop.set_lr(optimizer, lr)
        for q in (train_queue, val_queue):
            smp = getattr(q, 'sampler', None)
            if smp is not None and hasattr(smp, 'set_epoch'):
                smp.set_epoch(epoch)
        t0 = time.time()
        train_acc, train_obj = eval_loop.train(train_queue, model, criterion_smooth, optimizer, args, sync)
        logging.info('Train_acc: %f', train_acc)
        val1, val5, val_obj = eval_loop.validate(val_queue, model, criterion, args)
        logging.info('Val_acc_top1: %f', val1)
        logging.info('Val_acc_top5: %f', val5)
        logging.info('Epoch time: %ds.', time.time() - t0)
        is_best = val1 > best1
        if is_best:
            best1, best5 = val1, val5
        if rank == 0:       # tools/utils.py:112-117; keys carry the reference's DataParallel 'module.' prefix
            state = {'epoch': epoch + 1, 'state_dict': {'module.' + k: v for k, v in model.state_dict().items()},
                     'best_acc_top1': best1, 'best_acc_top5': best5, 'optimizer': optimizer.state_dict()}
            torch.save(state, os.path.join(args.save, 'checkpoint.pth.tar'))
            if is_best:
                torch.save(state, os.path.join(args.save, 'model_best.pth.tar'))


if __name__ == '__main__':
    main()
