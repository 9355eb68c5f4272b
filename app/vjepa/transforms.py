"""make_transforms with the reference's signature (app/vjepa/transforms.py:15).  The CPU video
augmentations (random-resized-crop, flip, RandAugment, erasing, normalise) are dataloader work outside
the accelerated path; synthetic clips are already N(0,1) and crop-sized, so this returns None
(= identity) and rejects configurations that would silently skip a requested augmentation."""


def make_transforms(random_horizontal_flip=True, random_resize_aspect_ratio=(3 / 4, 4 / 3),
                    random_resize_scale=(0.3, 1.0), reprob=0.0, auto_augment=False, motion_shift=False, crop_size=224,
                    normalize=((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
    if auto_augment or motion_shift or reprob > 0:
        raise NotImplementedError("auto_augment / motion_shift / random-erasing are CPU dataloader augmentations "
                                  "outside this package's scope")
    return None
