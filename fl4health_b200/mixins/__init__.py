from fl4health_b200.mixins.adaptive_drift_constrained import AdaptiveDriftConstrainedMixin, apply_adaptive_drift_to_client
from fl4health_b200.mixins.personalized import PersonalizedMode, make_it_personal

__all__ = ["AdaptiveDriftConstrainedMixin", "PersonalizedMode", "apply_adaptive_drift_to_client", "make_it_personal"]
