"""hvqm4_b200 -- B200-native HVQM4 1.3/1.5 picture decoder (host entropy stage in C,
pixel reconstruction in hand-written CUDA for sm_100a) behind the reference's
SDK-compatible C API.  See DESIGN.md and include/hvqm4.h."""
