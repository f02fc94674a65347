"""pnpvcve_b200 -- B200-native BAE+CAA enhancement hot path of PnP-VCVE."""
__version__ = "0.1.0"
