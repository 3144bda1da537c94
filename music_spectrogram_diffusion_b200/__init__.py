"""B200-native DDPM sampling hot path of music-spectrogram-diffusion."""
