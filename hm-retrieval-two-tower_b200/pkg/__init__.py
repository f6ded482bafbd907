"""B200-native drop-in for the reference's ``pkg`` namespace (hot path only)."""
import logging

logging.basicConfig(level=logging.INFO, format="%(asctime)s | %(name)s | %(levelname)s | %(message)s")
