"""B200-native mirror of the reference's ``quantization`` package (same module paths and names)."""
