"""swinvox_b200 -- B200-native (sm_100a) forward path of SwinVox multi-view reconstruction."""
__version__ = "0.1.0"
