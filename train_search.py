#!/usr/bin/env python
"""TF-NAS architecture search on B200 — same command line as the reference ``train_search.py``
(flags :29-66, epoch controller :155-315), running the supernet through tfnas_b200's CUDA kernels.

Extra, opt-in flags: ``--synthetic N`` (N synthetic batches per epoch instead of an image list;
there is no ImageNet in the build container) and torchrun env vars for data-parallel search
(one process per GPU; see tfnas_b200/parallel.py).  Checkpoints keep the reference format
``{'state_dict': max-width tensors with 'module.' prefix, 'mc_mask_dddict': masks}`` so
``parsing_model.py`` / ``train_eval.py`` of the reference consume them unchanged.
"""
import argparse
import copy
import logging
import os
import random
import sys
import time
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tfnas_b200 import config as cfg  # noqa: E402
from tfnas_b200 import elastic, model_search, parallel, parsing, search_loop  # noqa: E402
from tfnas_b200.lut import load_lut  # noqa: E402


def build_parser():
    p = argparse.ArgumentParser('searching TF-NAS')
    p.add_argument('--img_root', type=str, default=None, help='image root path (ImageNet train set)')
    p.add_argument('--train_list', type=str, default='./dataset/ImageNet-100-effb0_train_cls_ratio0.8.txt')
    p.add_argument('--val_list', type=str, default='./dataset/ImageNet-100-effb0_val_cls_ratio0.8.txt')
    p.add_argument('--lookup_path', type=str, default='./latency_pkl/latency_gpu.pkl', help='path of lookup table')
    p.add_argument('--save', type=str, default='./checkpoints', help='model and log saving path')
    p.add_argument('--print_freq', type=float, default=100)
    p.add_argument('--workers', type=int, default=4)
    p.add_argument('--epochs', type=int, default=90)
    p.add_argument('--batch_size', type=int, default=32)
    p.add_argument('--w_lr', type=float, default=0.025)
    p.add_argument('--w_mom', type=float, default=0.9)
    p.add_argument('--w_wd', type=float, default=1e-5)
    p.add_argument('--a_lr', type=float, default=0.01)
    p.add_argument('--a_wd', type=float, default=5e-4)
    p.add_argument('--a_beta1', type=float, default=0.5)
    p.add_argument('--a_beta2', type=float, default=0.999)
    p.add_argument('--grad_clip', type=float, default=5.0)
    p.add_argument('--T', type=float, default=5.0)
    p.add_argument('--T_decay', type=float, default=0.96)
    p.add_argument('--num_classes', type=int, default=100)
    p.add_argument('--seed', type=int, default=2)
    p.add_argument('--note', type=str, default='try')
    p.add_argument('--lambda_lat', type=float, default=0.1)
    p.add_argument('--target_lat', type=float, default=15.0)
    # additions
    p.add_argument('--synthetic', type=int, default=0, help='use N synthetic batches per epoch (no dataset needed)')
    p.add_argument('--warm_epochs', type=int, default=10, help='epochs of weight-only training (reference: 10)')
    p.add_argument('--image_size', type=int, default=224)
    return p


def cosine_lr_list(w_lr, epochs):
    """lr per epoch exactly as the reference obtains it (train_search.py:106-116): scheduler.get_lr()
    queried outside step(), including its chainable-form quirk."""
    opt = torch.optim.SGD([nn.Parameter(torch.zeros(1))], lr=w_lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, float(epochs))
    out = []
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for _ in range(epochs):
            out.append(sched.get_lr()[0])
            opt.step()
            sched.step()
    return out


class SyntheticQueue(object):
    """N batches of N(0,1) images / uniform labels, regenerated identically every epoch."""

    def __init__(self, n, bs, size, classes, seed):
        g = torch.Generator().manual_seed(seed)
        self.data = [(torch.randn(bs, 3, size, size, generator=g).pin_memory(),
                      torch.randint(0, classes, (bs,), generator=g).pin_memory()) for _ in range(min(n, 4))]
        self.n = n

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            yield self.data[i % len(self.data)]


def make_queues(args, rank, world):
    if args.synthetic > 0:
        return (SyntheticQueue(args.synthetic, args.batch_size, args.image_size, args.num_classes, args.seed + rank),
                SyntheticQueue(max(1, args.synthetic // 2), args.batch_size, args.image_size, args.num_classes,
                               1000 + args.seed + rank))
    if not args.img_root:
        raise SystemExit('--img_root is required unless --synthetic N is given')
    from tfnas_b200.data import make_imagenet_loaders
    return make_imagenet_loaders(args, rank, world)


def main(argv=None):
    args = build_parser().parse_args(argv)
    rank, local, world = parallel.init_from_env()
    # the kernels index pixels (N*H*W of the 112x112 first-stem plane at 224x224) with 23-bit fast division
    if args.batch_size * ((args.image_size + 1) // 2) ** 2 >= 1 << 23:
        raise SystemExit('--batch_size %d too large for one GPU at image size %d: N*H*W of the first-stem plane must stay '
                         'below 2^23 pixels (bs <= 668 at 224x224); use more ranks' % (args.batch_size, args.image_size))
    if not torch.cuda.is_available():
        logging.info('No GPU device available')
        sys.exit(1)
    torch.cuda.set_device(local)
    args.save = os.path.join(args.save, 'search-{}-{}'.format(time.strftime('%Y%m%d-%H%M%S'), args.note))
    if rank == 0:
        os.makedirs(args.save, exist_ok=True)
    fmt = '%(asctime)s %(message)s'
    logging.basicConfig(stream=sys.stdout, level=logging.INFO if rank == 0 else logging.WARNING, format=fmt,
                        datefmt='%m/%d %I:%M:%S %p')
    logging.getLogger().setLevel(logging.INFO if rank == 0 else logging.WARNING)     # basicConfig is a no-op under a host app
    if rank == 0:
        fh = logging.FileHandler(os.path.join(args.save, 'log.txt'))
        fh.setFormatter(logging.Formatter(fmt))
        logging.getLogger().addHandler(fh)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    random.seed(args.seed)
    if world > 1:
        model_search.seed_noise(args.seed)     # identical sampling on all ranks
    logging.info('args = %s', args)

    lat_lookup = load_lut(args.lookup_path)
    mc_mask_dddict = cfg.make_mc_mask_dddict()
    keys = cfg.lat_lookup_key_dddict
    mc_maxnum_dddict = cfg.get_mc_num_dddict(mc_mask_dddict, is_max=True)
    model = parallel.SearchParallel(model_search.Network(args.num_classes, mc_maxnum_dddict, lat_lookup)).cuda()
    logging.info('param size = %fMB', sum(np.prod(v.size()) for v in model.parameters()) / 1e6)
    path = lambda ep: os.path.join(args.save, 'searched_model_{:02}.pth.tar'.format(ep))
    state_dict = model.state_dict()
    if rank == 0:
        torch.save({'state_dict': state_dict, 'mc_mask_dddict': mc_mask_dddict}, path(0))
    lr_list = cosine_lr_list(args.w_lr, args.epochs)
    del model
    from tfnas_b200.step import FusedCrossEntropy
    criterion = FusedCrossEntropy().cuda()        # nn.CrossEntropyLoss() (reference :121) with its gradient in one launch
    train_queue, val_queue = make_queues(args, rank, world)
    sync = parallel.GradSync()

    for epoch in range(args.epochs):
        mc_num_dddict = cfg.get_mc_num_dddict(mc_mask_dddict)
        net = model_search.Network(args.num_classes, mc_num_dddict, lat_lookup)
        model = parallel.SearchParallel(net).cuda()
        net.set_temperature(args.T)
        elastic.load_from_master(model, state_dict, mc_mask_dddict)
        optimizer_w, optimizer_a = search_loop.make_optimizers(net, lr_list[epoch], args.w_mom, args.w_wd, args.a_lr,
                                                                args.a_beta1, args.a_beta2, args.a_wd)
        logging.info('Epoch: %d lr: %e T: %e', epoch, lr_list[epoch], args.T)
        for q in (train_queue, val_queue):      # distributed samplers reshuffle / repartition per epoch
            smp = getattr(q, 'sampler', None)
            if smp is not None and hasattr(smp, 'set_epoch'):
                smp.set_epoch(epoch)
        t0 = time.time()
        if epoch < args.warm_epochs:
            train_acc = search_loop.train_wo_arch(train_queue, model, criterion, optimizer_w, args, sync)
        else:
            train_acc = search_loop.train_w_arch(train_queue, val_queue, model, criterion, optimizer_w, optimizer_a,
                                                 args, sync)
            args.T *= args.T_decay
        logging.info('The current arch parameters are:')
        for param in net.log_alphas_parameters():
            logging.info(' '.join('{:.6f}'.format(p) for p in np.exp(param.detach().cpu().numpy())))
        for param in net.betas_parameters():
            logging.info(' '.join('{:.6f}'.format(p) for p in F.softmax(param.detach().cpu(), dim=-1).numpy()))
        logging.info('Train_acc %f', train_acc)
        logging.info('Epoch time: %ds', time.time() - t0)
        if args.epochs - epoch < 5:
            logging.info('Val_acc %f', search_loop.validate(val_queue, model, criterion, args))
        elastic.store_to_master(state_dict, model, mc_mask_dddict)
        if epoch >= args.warm_epochs:
            logging.info('Now shrinking or expanding the arch')
            op_w, depth_w = parsing.get_op_and_depth_weights(model)
            parsed_arch = parsing.parse_architecture(op_w, depth_w)
            mc_num_dddict, before, after = elastic.rescale_widths(parsed_arch, cfg.get_mc_num_dddict(mc_mask_dddict),
                                                                  mc_maxnum_dddict, keys, lat_lookup, args.target_lat)
            logging.info('Before, the current lat: {:.4f}, the target lat: {:.4f}'.format(before, args.target_lat))
            elastic.reselect_channels(mc_mask_dddict, mc_num_dddict, parsed_arch, state_dict)
            logging.info('After, the current lat: {:.4f}, the target lat: {:.4f}'.format(after, args.target_lat))
        if rank == 0:
            torch.save({'state_dict': state_dict, 'mc_mask_dddict': mc_mask_dddict}, path(epoch + 1))
        del model, net


if __name__ == '__main__':
    t0 = time.time()
    main()
    logging.info('Total searching time: %ds', time.time() - t0)
