"""Python wrappers over the sm_100a extension (filled in with the kernels)."""
