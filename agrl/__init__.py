"""agrl -- B200-native implementation of AGRL's test-time hot path (see agrl.pytorch_b200)."""
