"""Image-list datasets for real-data runs (reference dataset/dataset.py:33-49 + the transforms of
train_search.py:123-153 / train_eval.py:134-160).  Not used by the synthetic benchmarks."""
import os

import torch
import torch.utils.data as data

IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]


class ImageList(data.Dataset):
    """Lines of '<relative path> <label>'."""

    def __init__(self, root, list_path, transform=None):
        self.root, self.transform = root, transform
        self.samples = []
        with open(list_path) as f:
            for line in f:
                parts = line.strip().rsplit(' ', 1)
                if len(parts) == 2:
                    self.samples.append((parts[0], int(parts[1])))

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        from PIL import Image
        path, label = self.samples[i]
        with open(os.path.join(self.root, path), 'rb') as f:
            img = Image.open(f).convert('RGB')
        if self.transform is not None:
            img = self.transform(img)
        return img, label


def _transforms():
    import torchvision.transforms as T
    norm = T.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD)
    train_tf = T.Compose([T.RandomResizedCrop(224), T.RandomHorizontalFlip(),
                          T.ColorJitter(brightness=0.4, contrast=0.4, saturation=0.4, hue=0.2), T.ToTensor(), norm])
    val_tf = T.Compose([T.Resize(256), T.CenterCrop(224), T.ToTensor(), norm])
    return train_tf, val_tf


def make_eval_loaders(args, rank=0, world=1):
    """Loaders of the re-training run (train_eval.py:134-160); ``args.batch_size`` is global, each rank loads its share."""
    train_tf, val_tf = _transforms()
    sets = (ImageList(args.train_root, args.train_list, train_tf), ImageList(args.val_root, args.val_list, val_tf))
    out = []
    for ds, shuffle in zip(sets, (True, False)):
        sampler = data.distributed.DistributedSampler(ds, world, rank, shuffle=shuffle) if world > 1 else None
        out.append(data.DataLoader(ds, batch_size=args.batch_size // world, shuffle=shuffle and sampler is None,
                                   sampler=sampler, pin_memory=True, num_workers=max(1, args.workers // world)))
    return tuple(out)


def make_imagenet_loaders(args, rank=0, world=1):
    train_tf, val_tf = _transforms()
    train_set = ImageList(args.img_root, args.train_list, train_tf)
    val_set = ImageList(args.img_root, args.val_list, val_tf)

    def loader(ds):
        sampler = data.distributed.DistributedSampler(ds, world, rank, shuffle=True) if world > 1 else None
        return data.DataLoader(ds, batch_size=args.batch_size, shuffle=sampler is None, sampler=sampler,
                               pin_memory=True, num_workers=args.workers, drop_last=world > 1)
    return loader(train_set), loader(val_set)
