"""B200-native scoring-and-selection hot path of multi_view_active_learning."""
__version__ = "0.1.0"
