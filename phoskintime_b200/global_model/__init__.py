"""Global coupled kinase-TF-protein network path."""
from .network import GlobalSystem, synthetic_system, synthetic_loss_data  # noqa: F401
