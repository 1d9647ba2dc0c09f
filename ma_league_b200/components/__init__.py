from .episode_batch import EpisodeBatch
