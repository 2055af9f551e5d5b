"""puzzlelib_b200 -- B200-native (sm_100a) operator backend with the PuzzleLib Modules / Backend / GPUArray surface."""
__version__ = "0.1.0"
