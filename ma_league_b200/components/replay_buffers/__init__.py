from .replay_buffer import ReplayBuffer
