"""mlx_audio_b200 -- the B200-native drop-in for mlx-audio's speech-inference hot path.

Arrays are ``torch.Tensor`` on CUDA; every hot op is a hand-written sm_100a kernel reached through
the C ABI in ``include/b200audio.h`` (``libb200audio.so``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
