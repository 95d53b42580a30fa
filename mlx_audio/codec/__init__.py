from mlx_audio_b200.codec import SNAC, Mimi, MimiConfig, MimiStreamingDecoder, mimi_202407  # noqa: F401
