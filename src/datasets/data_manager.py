"""Data-loader factory with the reference's signature (src/datasets/data_manager.py:15-91).

Video decoding / augmentation (decord, PIL) is CPU dataloader work outside the accelerated hot path
(SURVEY.md section 2, rows 14-15).  What is provided here is the `synthetic` dataset the headline
metric is defined on (random N(0,1) clips, the distribution of normalised video) behind the same
`init_data(...) -> (loader, sampler)` contract, so `app.vjepa.train.main` runs end to end.
"""
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.distributed import DistributedSampler


class SyntheticVideoDataset(Dataset):
    """Items shaped like VideoDataset's: ([clip_0, ..., clip_{num_clips-1}], label, clip_indices)."""

    def __init__(self, length, num_frames, crop_size, num_clips=1, seed=0):
        self.length, self.num_frames, self.crop_size, self.num_clips, self.seed = length, num_frames, crop_size, num_clips, seed

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1000003 + index)
        clips = [torch.randn(3, self.num_frames, self.crop_size, self.crop_size, generator=g)
                 for _ in range(self.num_clips)]
        return clips, 0, [list(range(self.num_frames)) for _ in range(self.num_clips)]


class SyntheticUint8VideoDataset(Dataset):
    """Decoded-video stand-in: uint8 frames [T, H, W, 3] (H, W a bit larger than the crop, like decord output after the
    short-side resize) handed to `transform` exactly as VideoDataset hands its buffer to the reference transform
    (src/datasets/video_dataset.py:157-163).  With the GPU input pipeline the transform returns ClipTickets."""

    def __init__(self, length, num_frames, frame_hw, transform, num_clips=1, seed=0):
        self.length, self.num_frames, self.frame_hw, self.transform = length, num_frames, frame_hw, transform
        self.num_clips, self.seed = num_clips, seed

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        g = torch.Generator().manual_seed(self.seed * 1000003 + index)
        H, W = self.frame_hw
        clips = []
        for _ in range(self.num_clips):
            buf = torch.randint(0, 256, (self.num_frames, H, W, 3), dtype=torch.uint8, generator=g).numpy()
            clips.append(self.transform(buf) if self.transform is not None else torch.from_numpy(buf))
        return clips, 0, [list(range(self.num_frames)) for _ in range(self.num_clips)]


def init_data(batch_size, transform=None, shared_transform=None, data='ImageNet', collator=None, pin_mem=True,
              num_workers=8, world_size=1, rank=0, root_path=None, image_folder=None, training=True, copy_data=False,
              drop_last=True, tokenize_txt=True, subset_file=None, clip_len=8, frame_sample_rate=2, duration=None,
              num_clips=1, random_clip_sampling=True, allow_clip_overlap=False, filter_short_videos=False,
              filter_long_videos=int(1e9), decode_one_clip=True, datasets_weights=None, persistent_workers=False,
              repeat_wds=False, ipe=300, log_dir=None, crop_size=224, synthetic_length=None):
    kind = str(data).lower()
    if kind == 'synthetic_uint8':
        # exercises the GPU input pipeline end to end: uint8 frames + host-side crop / flip decisions -> ClipTickets
        length = synthetic_length or batch_size * world_size * ipe
        hw = (int(crop_size * 8 / 7) // 2 * 2, int(crop_size * 4 / 3) // 2 * 2)
        dataset = SyntheticUint8VideoDataset(length, clip_len, hw, transform, num_clips=num_clips, seed=rank)
        sampler = DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=True)
        loader = DataLoader(dataset, collate_fn=collator, sampler=sampler, batch_size=batch_size, drop_last=drop_last,
                            pin_memory=False, num_workers=num_workers, persistent_workers=num_workers > 0)
        return loader, sampler
    if kind != 'synthetic':
        raise NotImplementedError(
            f"dataset_type={data!r}: video/image decoding pipelines (decord / PIL, src/datasets/video_dataset.py in the "
            "reference) are outside this package's scope; use dataset_type: synthetic or plug your own "
            "torch DataLoader whose collate_fn is the mask collator")
    length = synthetic_length or batch_size * world_size * ipe
    dataset = SyntheticVideoDataset(length, clip_len, crop_size, num_clips=num_clips, seed=rank)
    sampler = DistributedSampler(dataset, num_replicas=world_size, rank=rank, shuffle=True)
    loader = DataLoader(dataset, collate_fn=collator, sampler=sampler, batch_size=batch_size, drop_last=drop_last,
                        pin_memory=pin_mem, num_workers=num_workers, persistent_workers=num_workers > 0)
    return loader, sampler
