"""B200-native batched Smith-Waterman engine behind MegaPath-Nano's ssw.h C ABI."""
