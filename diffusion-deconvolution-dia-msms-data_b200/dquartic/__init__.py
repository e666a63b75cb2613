"""dquartic (B200-native build).  Same package layout as the reference (`dquartic.model`, `dquartic.utils`,
`dquartic.cli`); the hot path runs in hand-written sm_100a kernels behind include/dquartic_b200.h."""
__version__ = "0.1.0+b200"
