"""B200-native MultiSWAG posterior-predictive / SWAG-training hot path of bnn_chaos_model."""
__version__ = "0.1.0"
