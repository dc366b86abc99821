"""B200-native wavelet-monodepth decoder hot path."""
