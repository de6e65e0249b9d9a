"""B200-native 802.11a/g OFDM baseband: drop-in for the `wifi_phy_hier` block of
OedonLestrange42/GNURadio-WiFI-ImageTransfer (gnu_radio/wifi_phy_hier.grc).

The directory name carries a hyphen (it mirrors the reference repository's name), so it is
imported through the `wifi_b200` shim at the repository root or with importlib:
    import importlib; pkg = importlib.import_module("gnuradio-wifi-imagetransfer_b200")
"""
from . import build, wifi_b200  # noqa: F401
from .wifi_b200 import Handle, WifiB200Error, FRAME_DTYPE, ENCODINGS, EQUALIZERS  # noqa: F401
from .wifi_phy_hier import wifi_phy_hier, mac  # noqa: F401
from . import sharding, loopback_runner, ota_runners, pcap, featuremap  # noqa: F401
