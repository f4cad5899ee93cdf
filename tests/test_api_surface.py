"""Public API parity with the reference: every module / class / method name a
dist-keras user relies on must exist, with the reference's constructor defaults."""
import importlib
import inspect

import pytest

SURFACE = {
    "trainers": {
        "Trainer": ["set_max_prefetch", "set_model", "record_training_start", "record_training_end",
                    "get_training_time", "get_history", "get_averaged_history", "get_executor_history", "train",
                    "serialize"],
        "SingleTrainer": ["allocate_worker", "train"],
        "AveragingTrainer": ["average_models", "allocate_worker", "train"],
        "EnsembleTrainer": ["allocate_worker", "train"],
        "DistributedTrainer": ["set_minibatch_size", "get_minibatch_size", "get_features_column", "get_label_column",
                               "get_learning_rate", "set_learning_rate", "set_num_epoch", "get_num_epoch",
                               "allocate_worker", "set_master", "determine_new_master", "allocate_parameter_server",
                               "set_num_workers", "get_num_workers", "num_updates", "service", "stop_service",
                               "start_service", "train"],
        "AsynchronousDistributedTrainer": ["set_parallelism_factor", "get_parallelism_factor", "train"],
        "AEASGD": [], "DOWNPOUR": [], "EAMSGD": [], "ADAG": [], "DynSGD": [], "Experimental": [],
    },
    "workers": {
        "Worker": ["set_max_prefetch", "set_learning_rate", "get_learning_rate", "set_worker_id", "get_worker_id",
                   "prepare_model", "get_next_minibatch", "start_prefetching_thread", "prefetching", "optimize",
                   "train"],
        "SequentialWorker": [],
        "NetworkWorker": ["connect", "pull", "commit", "set_tcp_no_delay", "tcp_no_delay", "get_master_host",
                          "get_master_port", "add_history", "train"],
        "ADAGWorker": [], "DOWNPOURWorker": [], "AEASGDWorker": [], "EAMSGDWorker": [], "DynSGDWorker": [],
        "ExperimentalWorker": [],
    },
    "parameter_servers": {
        "ParameterServer": ["initialize", "start", "run", "stop", "get_model", "next_update", "reset_update_counter",
                            "get_num_updates"],
        "SocketParameterServer": ["initialize", "handle_commit", "handle_pull", "cancel_accept", "handle_connection",
                                  "start", "run", "stop", "finalize", "cleanup_connections"],
        "DeltaParameterServer": [], "ADAGParameterServer": [], "DynSGDParameterServer": [],
        "ExperimentalParameterServer": [],
    },
    "networking": {n: None for n in ["determine_host_address", "recvall", "recv_data", "send_data", "connect"]},
    "utils": {n: None for n in ["get_os_username", "set_keras_base_directory", "to_one_hot_encoded_dense",
                                "new_dataframe_row", "json_to_dataframe_row", "pickle_object", "unpickle_object",
                                "serialize_keras_model", "deserialize_keras_model", "history_executors_average",
                                "history_executor", "uniform_weights", "shuffle", "precache"]},
    "transformers": {"Transformer": ["transform"], "MinMaxTransformer": [], "BinaryLabelTransformer": [],
                     "StandardTransformer": [], "DenseTransformer": [], "ReshapeTransformer": [],
                     "OneHotTransformer": [], "LabelIndexTransformer": []},
    "predictors": {"Predictor": ["predict"], "ModelPredictor": []},
    "evaluators": {"Evaluator": ["evaluate"], "AccuracyEvaluator": []},
    "schemes": {"Scheme": [], "Emperor": ["optimize"]},
    "job_deployment": {"Job": ["set_num_executors", "set_num_processes", "get_trained_model", "get_history",
                               "is_finished", "destroy_remote_job", "start", "wait_completion", "cancel", "send",
                               "run"],
                       "Punchcard": [], "PunchcardJob": []},
}

# constructor defaults quoted in the reference (distkeras/trainers.py:672-893, transformers.py:313)
DEFAULTS = {
    ("trainers", "SingleTrainer"): {"num_epoch": 1, "batch_size": 32, "features_col": "features",
                                    "label_col": "label"},
    ("trainers", "AveragingTrainer"): {"num_workers": 2, "batch_size": 32, "num_epoch": 1},
    ("trainers", "EnsembleTrainer"): {"num_ensembles": 2, "batch_size": 32},
    ("trainers", "DistributedTrainer"): {"num_workers": 2, "batch_size": 32, "master_port": 5000, "num_epoch": 1},
    ("trainers", "AEASGD"): {"communication_window": 32, "rho": 5.0, "learning_rate": 0.1},
    ("trainers", "DOWNPOUR"): {"communication_window": 5},
    ("trainers", "EAMSGD"): {"communication_window": 32, "rho": 5.0, "learning_rate": 0.1, "momentum": 0.9},
    ("trainers", "ADAG"): {"communication_window": 12},
    ("trainers", "DynSGD"): {"communication_window": 5},
    ("trainers", "Experimental"): {"communication_window": 5, "learning_rate": 1.0},
    ("transformers", "LabelIndexTransformer"): {"input_col": "prediction", "output_col": "prediction_index",
                                                "default_index": 0, "activation_threshold": 0.55},
    ("transformers", "StandardTransformer"): {"suffix": "_normalized"},
    ("transformers", "MinMaxTransformer"): {"is_vector": True},
    ("predictors", "ModelPredictor"): {"features_col": "features", "output_col": "prediction"},
    ("evaluators", "Evaluator"): {"label_col": "label", "prediction_col": "prediction"},
    ("schemes", "Emperor"): {"num_epoch": 15, "evaluation_frequency": 5, "loss_threshold": 0.005},
    ("job_deployment", "Punchcard"): {"secrets_path": "secrets.json", "port": 80},
}


@pytest.mark.parametrize("module", sorted(SURFACE))
def test_names_exist(module):
    mod = importlib.import_module("distkeras_b200." + module)
    missing = []
    for name, methods in SURFACE[module].items():
        if not hasattr(mod, name):
            missing.append(name)
            continue
        for meth in methods or []:
            if not callable(getattr(getattr(mod, name), meth, None)):
                missing.append(f"{name}.{meth}")
    assert not missing, f"{module}: missing {missing}"


@pytest.mark.parametrize("key", sorted(DEFAULTS))
def test_constructor_defaults(key):
    module, cls = key
    obj = getattr(importlib.import_module("distkeras_b200." + module), cls)
    params = inspect.signature(obj.__init__).parameters
    for arg, want in DEFAULTS[key].items():
        assert arg in params, f"{cls}.__init__ has no parameter {arg!r}"
        assert params[arg].default == want, f"{cls}({arg}=...) default {params[arg].default!r} != {want!r}"


def test_alias_package_importable():
    """The hyphenated directory name in the task statement resolves to the same package."""
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert os.path.realpath(os.path.join(root, "dist-keras_b200")) == os.path.realpath(
        os.path.join(root, "distkeras_b200"))


def test_cli_info_reports_environment():
    import json
    import subprocess
    import sys

    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "distkeras_b200", "info"], cwd=root, capture_output=True, text=True,
                       timeout=300, env=dict(__import__("os").environ, DK_STRICT="1"))
    assert r.returncode == 0, r.stderr[-1000:]
    d = json.loads(r.stdout)
    assert "thread" in d["backends"] and d["switches"].get("DK_STRICT") == "1"
