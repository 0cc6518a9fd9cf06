"""rlipv2_b200 - B200-native (sm_100a) hot path of RLIPv2-ParSeDA behind the reference's own
operator surface.  See DESIGN.md for the path and INTEGRATION.md for the drop-in binding."""
__version__ = "0.1.0"
