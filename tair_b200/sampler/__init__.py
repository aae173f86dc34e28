from .spaced_sampler import Sampler, SpacedSampler, space_timesteps  # noqa: F401
