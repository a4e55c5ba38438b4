"""elemental_b200 -- B200-native dense level-3 hot path behind Elemental's API."""
