"""B200-native stereo-MSCKF hot path (drop-in for mfkiwl/msckf_stereo_c's per-frame path)."""
