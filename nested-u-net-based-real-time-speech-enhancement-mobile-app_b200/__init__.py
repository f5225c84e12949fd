"""B200-native NUNet-TLS / NUNet-TLS-LSTM inference path (host side; the compute is csrc/*.cu)."""
