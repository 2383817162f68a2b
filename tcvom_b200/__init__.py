"""tcvom_b200 -- B200-native (sm_100a) implementation of TCVOM's GCA+TAM frame-window hot path."""
__version__ = "0.1.0"
