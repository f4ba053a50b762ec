from .train_flows import get_params, predict, set_params, train, importance_weights, svi_importance  # noqa: F401
