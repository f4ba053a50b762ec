from .train_flows import get_params, predict, set_params, importance_weights, svi_importance  # noqa: F401
