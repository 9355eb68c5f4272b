"""src.models.utils.patch_embed -> jepa_b200.models."""
from jepa_b200.models import PatchEmbed, PatchEmbed3D  # noqa: F401
