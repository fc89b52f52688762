"""ganon_b200 -- the ganon-classify hot path on NVIDIA B200."""
__all__ = ["classify", "cli", "formats"]
