"""distkeras_b200: a B200-native asynchronous parameter-server training framework with the
capabilities and API surface of cerndb/dist-keras (trainers / workers / parameter_servers /
networking / utils / transformers / predictors / evaluators / schemes / job_deployment)."""
from . import (data, evaluators, job_deployment, models, networking, ops, parameter_servers, predictors, schemes,
               trainers, transformers, utils, workers)
from .data import Dataset, Row
from .models import Sequential, model_from_json

__version__ = "0.1.0"
__all__ = ["data", "evaluators", "job_deployment", "models", "networking", "ops", "parameter_servers",
           "predictors", "schemes", "trainers", "transformers", "utils", "workers", "Dataset", "Row",
           "Sequential", "model_from_json"]
