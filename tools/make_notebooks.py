#!/usr/bin/env python
"""Writes the example notebooks that mirror the reference's per-topic notebooks one to one
(``/root/reference/examples/*.ipynb``), on this framework's API and synthetic stand-in data (no network here).

    python tools/make_notebooks.py        # (re)writes examples/<name>.ipynb, unexecuted

The cells are plain strings below; ``tests/test_examples.py`` runs every notebook's code top to bottom at toy sizes.
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HEADER = 'import os, sys, tempfile, time\nsys.path.insert(0, os.path.abspath(".."))          # run from examples/\nimport numpy as np, torch\n'

NOTEBOOKS = {}

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["mnist_preprocessing"] = [
    ("md", "# MNIST preprocessing\n\nCounterpart of the reference's `mnist_preprocessing.ipynb`: read the raw table, assemble the pixel "
           "columns into one vector, one-hot the label, scale the pixels to [0, 1], reshape a copy for the convolutional "
           "model, enlarge the training set by repeated `unionAll`, and store everything as Parquet. The real CSV is not "
           "reachable from this sandbox, so `synthetic_mnist` writes a CSV of the same layout first."),
    ("code", HEADER + "from distkeras_b200.data import Dataset, synthetic_mnist\n"
             "from distkeras_b200.transformers import MinMaxTransformer, OneHotTransformer, ReshapeTransformer\n"
             "tmp = tempfile.mkdtemp(prefix='dk_mnist_')"),
    ("md", "## A CSV with the layout of `mnist_train.csv` (label, then 784 pixel columns)"),
    ("code", "raw = synthetic_mnist(4000, as_uint8=False)\n"
             "table = np.concatenate([raw['label'].numpy()[:, None].astype(np.float32), raw['features'].numpy()], axis=1)\n"
             "header = ','.join(['label'] + [f'p{i}' for i in range(784)])\n"
             "csv_path = os.path.join(tmp, 'mnist_train.csv')\n"
             "np.savetxt(csv_path, table, delimiter=',', header=header, comments='', fmt='%g')\n"
             "print(os.path.getsize(csv_path) // 1024, 'KiB')"),
    ("md", "## Read + assemble\n`Dataset.from_csv` plays the role of the CSV reader plus `VectorAssembler`: every non-label column "
           "lands in one `features` vector column."),
    ("code", "t0 = time.time()\ndataset = Dataset.from_csv(csv_path, label_col='label', num_partitions=4)\ndataset.printSchema()\nprint(dataset.count(), 'rows')"),
    ("md", "## Label and feature transformations"),
    ("code", "dataset = OneHotTransformer(10, input_col='label', output_col='label_encoded').transform(dataset)\n"
             "dataset = MinMaxTransformer(o_min=0.0, o_max=255.0, n_min=0.0, n_max=1.0, input_col='features',\n"
             "                            output_col='features_normalized').transform(dataset)\n"
             "dataset = ReshapeTransformer('features_normalized', 'matrix', (28, 28, 1)).transform(dataset)\n"
             "print(dataset.first()['label_encoded'], tuple(dataset['matrix'].shape))"),
    ("md", "## Enlarging the training set\nThe reference builds `mnist_train_big.parquet` by unioning the table with itself ten "
           "times to have something that takes a while on a cluster; three doublings here."),
    ("code", "big = dataset\nfor _ in range(3):\n    big = big.unionAll(big)\nprint(big.count(), 'rows after enlarging')"),
    ("md", "## Shuffle and store"),
    ("code", "from distkeras_b200.utils import shuffle\n"
             "dataset = shuffle(dataset, seed=0)\n"
             "out = os.path.join(tmp, 'mnist_train.parquet')\n"
             "dataset.select('features_normalized', 'label', 'label_encoded').to_parquet(out)\n"
             "back = Dataset.from_parquet(out)\n"
             "print(back.columns, back.count(), 'rows;', round(time.time() - t0, 2), 's')"),
]

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["mnist_analysis"] = [
    ("md", "# MNIST analysis\n\nCounterpart of the reference's `mnist_analysis.ipynb`: the 784-1000-200-10 multilayer perceptron "
           "(its cell 11, the same network `bench.py` times) trained with ADAG at the notebook's tiny mini-batch, then the "
           "predict -> label-index -> accuracy pipeline. On a B200 the trainer runs the captured-graph fabric backend; on "
           "a CPU it falls back to worker threads."),
    ("code", HEADER + "from distkeras_b200.data import synthetic_mnist\n"
             "from distkeras_b200.models import Sequential, Dense, Dropout, Activation\n"
             "from distkeras_b200.trainers import ADAG, DOWNPOUR\n"
             "from distkeras_b200.predictors import ModelPredictor\n"
             "from distkeras_b200.transformers import LabelIndexTransformer, MinMaxTransformer, OneHotTransformer\n"
             "from distkeras_b200.evaluators import AccuracyEvaluator\n"
             "num_workers = max(1, torch.cuda.device_count())"),
    ("md", "## Data\nPixels scaled to [0, 1] (`features_normalized_dense` in the reference), label one-hot encoded."),
    ("code", "raw = synthetic_mnist(60000, as_uint8=False)\n"
             "raw = MinMaxTransformer(0.0, 255.0, 0.0, 1.0, 'features', 'features_normalized_dense').transform(raw)\n"
             "raw = OneHotTransformer(10, 'label', 'label_encoded').transform(raw)\n"
             "training_set, test_set = raw.randomSplit([0.85, 0.15], seed=0)\n"
             "training_set = training_set.repartition(num_workers).cache()\n"
             "print(training_set.count(), 'training rows,', test_set.count(), 'test rows')"),
    ("md", "## Model"),
    ("code", "mlp = Sequential()\n"
             "mlp.add(Dense(1000, input_shape=(784,)))\nmlp.add(Activation('relu'))\nmlp.add(Dropout(0.2))\n"
             "mlp.add(Dense(200))\nmlp.add(Activation('relu'))\nmlp.add(Dropout(0.2))\n"
             "mlp.add(Dense(10))\nmlp.add(Activation('softmax'))\nprint(mlp.summary())"),
    ("md", "## Evaluation helper\nPrediction column -> class index (`LabelIndexTransformer`) -> share of matches."),
    ("code", "def evaluate_accuracy(model, dataset, features='features_normalized_dense'):\n"
             "    predicted = ModelPredictor(keras_model=model, features_col=features).predict(dataset)\n"
             "    indexed = LabelIndexTransformer(output_dim=10).transform(predicted)\n"
             "    return AccuracyEvaluator(prediction_col='prediction_index', label_col='label').evaluate(indexed)"),
    ("md", "## ADAG, mini-batch 4, window 5 (the reference's setting)"),
    ("code", "trainer = ADAG(keras_model=mlp, worker_optimizer='adam', loss='categorical_crossentropy', num_workers=num_workers,\n"
             "               batch_size=4, communication_window=5, num_epoch=1,\n"
             "               features_col='features_normalized_dense', label_col='label_encoded')\n"
             "trained = trainer.train(training_set)\n"
             "print('training time %.2f s, %d center updates' % (trainer.get_training_time(), trainer.num_updates()))\n"
             "print('accuracy', evaluate_accuracy(trained, test_set))"),
    ("md", "## The loss over time, averaged over the workers"),
    ("code", "avg = trainer.get_averaged_history()\nprint('first / last averaged loss: %.3f -> %.3f' % (avg[0][0], avg[-1][0]))"),
    ("md", "## The same network at mini-batch 64 with DOWNPOUR"),
    ("code", "trainer = DOWNPOUR(keras_model=mlp, worker_optimizer='adam', loss='categorical_crossentropy', num_workers=num_workers,\n"
             "                   batch_size=64, communication_window=5, num_epoch=1,\n"
             "                   features_col='features_normalized_dense', label_col='label_encoded')\n"
             "trained = trainer.train(training_set)\n"
             "print('training time %.2f s' % trainer.get_training_time(), 'accuracy', evaluate_accuracy(trained, test_set))"),
]

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["example_0_data_preprocessing"] = [
    ("md", "# Example 0: data preprocessing\n\nCounterpart of the reference's `example_0_data_preprocessing.ipynb`: the ATLAS "
           "Higgs CSV -> drop the bookkeeping columns -> assemble the features -> standardise -> index and one-hot the "
           "string label -> shuffle -> Parquet. `examples/data/make_sample_data.py` writes a CSV with the challenge's "
           "column layout."),
    ("code", HEADER + "import subprocess\n"
             "from distkeras_b200.data import Dataset\n"
             "from distkeras_b200.transformers import OneHotTransformer, StandardTransformer\n"
             "from distkeras_b200.utils import shuffle\n"
             "tmp = tempfile.mkdtemp(prefix='dk_higgs_')\n"
             "subprocess.check_call([sys.executable, os.path.join('data', 'make_sample_data.py'), '--rows', '20000', '--out', tmp])"),
    ("md", "## Reading\n`EventId` and `Weight` are not features; `Label` is the string `s` / `b`."),
    ("code", "time_start = time.time()\n"
             "dataset = Dataset.from_csv(os.path.join(tmp, 'atlas_higgs_sample.csv'), label_col='Label',\n"
             "                           drop_cols=('EventId', 'Weight'), label_map={'b': 0, 's': 1}, num_partitions=4)\n"
             "dataset = dataset.withColumnRenamed('Label', 'label_index')\n"
             "dataset.printSchema()\nprint(dataset.take(1)[0]['features'][:5])"),
    ("md", "## Feature normalisation (zero mean, unit variance per feature)"),
    ("code", "dataset = StandardTransformer(['features']).transform(dataset)\n"
             "f = dataset['features_normalized']\nprint('mean %.3g, std %.3g' % (float(f.mean()), float(f.std())))"),
    ("md", "## Label transformation\nThe loss wants a vector with one entry per output neuron."),
    ("code", "dataset = OneHotTransformer(2, input_col='label_index', output_col='label').transform(dataset)\n"
             "print(dataset.select('label_index', 'label').take(3))"),
    ("md", "## Shuffle and save"),
    ("code", "dataset = shuffle(dataset, seed=0)\n"
             "out = os.path.join(tmp, 'processed.parquet')\n"
             "dataset.to_parquet(out)\n"
             "print('total time %.2f s' % (time.time() - time_start), '->', out)"),
]

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["example_1_analysis"] = [
    ("md", "# Example 1: model development and evaluation\n\nCounterpart of the reference's `example_1_analysis.ipynb`: a "
           "500-500-500 perceptron on the preprocessed Higgs table, trained once with `SingleTrainer` and then with the "
           "asynchronous optimizers, compared on F1 and training time."),
    ("code", HEADER + "from distkeras_b200.data import synthetic_higgs\n"
             "from distkeras_b200.models import Sequential, Dense, Dropout, Activation\n"
             "from distkeras_b200.trainers import ADAG, AEASGD, DOWNPOUR, SingleTrainer\n"
             "from distkeras_b200.predictors import ModelPredictor\n"
             "from distkeras_b200.transformers import LabelIndexTransformer, OneHotTransformer, StandardTransformer\n"
             "from distkeras_b200.evaluators import F1Evaluator\n"
             "num_workers = max(2, torch.cuda.device_count())"),
    ("md", "## Data\n(What `example_0_data_preprocessing.ipynb` stores; generated here so that the notebook stands alone.)"),
    ("code", "raw = synthetic_higgs(200000)\n"
             "raw = StandardTransformer(['features']).transform(raw)\n"
             "raw = OneHotTransformer(2, 'label', 'label_encoded').transform(raw).withColumnRenamed('label', 'label_index')\n"
             "nb_features, nb_classes = raw['features_normalized'].shape[1], 2\n"
             "training_set, test_set = raw.randomSplit([0.7, 0.3], seed=0)\n"
             "training_set = training_set.repartition(num_workers).cache()\n"
             "print(nb_features, 'features;', training_set.count(), 'training rows,', test_set.count(), 'test rows')"),
    ("md", "## Model"),
    ("code", "def build():\n"
             "    model = Sequential()\n"
             "    model.add(Dense(500, input_shape=(nb_features,)))\n    model.add(Activation('relu'))\n    model.add(Dropout(0.4))\n"
             "    model.add(Dense(500))\n    model.add(Activation('relu'))\n    model.add(Dropout(0.6))\n"
             "    model.add(Dense(500))\n    model.add(Activation('relu'))\n"
             "    model.add(Dense(nb_classes))\n    model.add(Activation('softmax'))\n    return model\n"
             "optimizer, loss = 'adagrad', 'categorical_crossentropy'\nprint(build().summary())"),
    ("md", "## Evaluation: F1 of the signal class"),
    ("code", "def evaluate(model):\n"
             "    predicted = ModelPredictor(keras_model=model, features_col='features_normalized').predict(test_set)\n"
             "    indexed = LabelIndexTransformer(output_dim=nb_classes).transform(predicted)\n"
             "    return F1Evaluator(label_col='label_index', prediction_col='prediction_index').evaluate(indexed)\n"
             "results, time_spent = {}, {}"),
    ("md", "## Single trainer"),
    ("code", "trainer = SingleTrainer(keras_model=build(), loss=loss, worker_optimizer=optimizer, features_col='features_normalized',\n"
             "                        label_col='label_encoded', num_epoch=1, batch_size=64)\n"
             "model = trainer.train(training_set)\n"
             "results['single'], time_spent['single'] = evaluate(model), trainer.get_training_time()\nprint(results, time_spent)"),
    ("md", "## Asynchronous EASGD, DOWNPOUR, ADAG"),
    ("code", "runs = {\n"
             "    'aeasgd': lambda: AEASGD(keras_model=build(), worker_optimizer=optimizer, loss=loss, num_workers=num_workers, batch_size=64,\n"
             "                             features_col='features_normalized', label_col='label_encoded', num_epoch=1,\n"
             "                             communication_window=32, rho=5.0, learning_rate=0.1),\n"
             "    'downpour': lambda: DOWNPOUR(keras_model=build(), worker_optimizer=optimizer, loss=loss, num_workers=num_workers,\n"
             "                                 batch_size=64, communication_window=5, num_epoch=1,\n"
             "                                 features_col='features_normalized', label_col='label_encoded'),\n"
             "    'adag': lambda: ADAG(keras_model=build(), worker_optimizer=optimizer, loss=loss, num_workers=num_workers, batch_size=64,\n"
             "                         communication_window=12, num_epoch=1, features_col='features_normalized',\n"
             "                         label_col='label_encoded'),\n"
             "}\n"
             "for name, make in runs.items():\n"
             "    trainer = make()\n"
             "    model = trainer.train(training_set)\n"
             "    results[name], time_spent[name] = evaluate(model), trainer.get_training_time()\n"
             "    print('%-9s F1 %.3f  %.2f s  %d center updates' % (name, results[name], time_spent[name], trainer.num_updates()))"),
    ("md", "## Results\n(The reference draws two bar charts; a table says the same without a plotting dependency.)"),
    ("code", "print('%-10s %8s %10s' % ('optimizer', 'F1', 'seconds'))\n"
             "for name in results:\n    print('%-10s %8.3f %10.2f' % (name, results[name], time_spent[name]))"),
]

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["cifar-10-preprocessing"] = [
    ("md", "# CIFAR-10 preprocessing\n\nCounterpart of the reference's `cifar-10-preprocessing.ipynb`: the pickled python batches "
           "(`data_batch_1..5`, `test_batch`: `{'data': uint8 [10000, 3072], 'labels': [...]}`, channel-major rows) -> CSV with "
           "one column per pixel -> vector column -> one-hot label -> pixels scaled to [0, 1] -> Parquet. The download is "
           "replaced by batches of the same format written from `synthetic_cifar10`."),
    ("code", HEADER + "import csv, pickle\n"
             "from distkeras_b200.data import Dataset, synthetic_cifar10\n"
             "from distkeras_b200.transformers import MinMaxTransformer, OneHotTransformer, ReshapeTransformer\n"
             "tmp = tempfile.mkdtemp(prefix='dk_cifar_')"),
    ("md", "## Batches in the format of `cifar-10-batches-py`"),
    ("code", "def write_batch(path, ds):\n"
             "    nhwc = ds['features'].numpy()                         # [n, 32, 32, 3]\n"
             "    planar = nhwc.transpose(0, 3, 1, 2).reshape(len(nhwc), -1)   # all red, then green, then blue\n"
             "    with open(path, 'wb') as f:\n"
             "        pickle.dump({'data': planar, 'labels': ds['label'].tolist()}, f)\n"
             "for i in range(1, 3):\n    write_batch(os.path.join(tmp, f'data_batch_{i}'), synthetic_cifar10(1000, seed=i))\n"
             "write_batch(os.path.join(tmp, 'test_batch'), synthetic_cifar10(500, seed=9))"),
    ("md", "## Load the batches and write the per-pixel CSV"),
    ("code", "def load(paths):\n"
             "    xs, ys = [], []\n"
             "    for p in paths:\n"
             "        with open(p, 'rb') as f:\n            d = pickle.load(f)\n"
             "        xs.append(np.asarray(d['data']))\n        ys += list(d['labels'])\n"
             "    return np.concatenate(xs), np.asarray(ys)\n"
             "columns = ['label'] + [f'p_{i}_{c}' for c in 'rgb' for i in range(1024)]\n"
             "def save(path, x, y):\n"
             "    with open(path, 'w', newline='') as f:\n"
             "        w = csv.writer(f)\n        w.writerow(columns)\n"
             "        for row, label in zip(x, y):\n            w.writerow([int(label)] + row.tolist())\n"
             "x_train, y_train = load([os.path.join(tmp, f'data_batch_{i}') for i in range(1, 3)])\n"
             "x_test, y_test = load([os.path.join(tmp, 'test_batch')])\n"
             "save(os.path.join(tmp, 'cifar-10-training.csv'), x_train, y_train)\n"
             "save(os.path.join(tmp, 'cifar-10-test.csv'), x_test, y_test)\n"
             "print(x_train.shape, x_test.shape)"),
    ("md", "## Vector column, one-hot label, normalised pixels"),
    ("code", "def prepare(path):\n"
             "    ds = Dataset.from_csv(path, label_col='label', num_partitions=4)\n"
             "    ds = OneHotTransformer(10, input_col='label', output_col='label_encoded').transform(ds)\n"
             "    ds = MinMaxTransformer(o_min=0.0, o_max=255.0, n_min=0.0, n_max=1.0, input_col='features',\n"
             "                           output_col='features_normalized').transform(ds)\n"
             "    return ds\n"
             "train, test = prepare(os.path.join(tmp, 'cifar-10-training.csv')), prepare(os.path.join(tmp, 'cifar-10-test.csv'))\n"
             "train.printSchema()\nprint(train.count(), test.count())"),
    ("md", "## Channel-last images for the convolutional model\nRows are planar (RRR..GGG..BBB); `Conv2D` wants `[32, 32, 3]`."),
    ("code", "def to_nhwc(ds):\n"
             "    x = ds['features_normalized'].reshape(-1, 3, 32, 32).permute(0, 2, 3, 1).contiguous()\n"
             "    return ds.with_column('image', x)\n"
             "train, test = to_nhwc(train), to_nhwc(test)\nprint(tuple(train['image'].shape))"),
    ("md", "## Parquet"),
    ("code", "train.select('features_normalized', 'label', 'label_encoded').to_parquet(os.path.join(tmp, 'cifar-10-train-preprocessed.parquet'))\n"
             "test.select('features_normalized', 'label', 'label_encoded').to_parquet(os.path.join(tmp, 'cifar-10-test-preprocessed.parquet'))\n"
             "print(sorted(os.listdir(tmp)))"),
    ("md", "## One epoch of the CIFAR-10 CNN on the result"),
    ("code", "from distkeras_b200.models import cifar10_cnn\nfrom distkeras_b200.trainers import SingleTrainer\n"
             "t = SingleTrainer(cifar10_cnn(seed=0), 'adam', 'categorical_crossentropy', features_col='image', label_col='label_encoded',\n"
             "                  batch_size=50, num_epoch=1)\n"
             "t.train(train)\nh = t.get_history()\nprint('loss %.3f -> %.3f' % (h[0]['history'][0], h[-1]['history'][0]))"),
]

# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["distributed_numpy_parsing"] = [
    ("md", "# Distributed numpy parsing\n\nCounterpart of the reference's `distributed_numpy_parsing.ipynb`: a directory of `.npy` "
           "files holding structured arrays (one record per particle, named fields) -> a table of paths -> every partition "
           "parses its own files -> one `Dataset` with one column per field, stored as Parquet. HDFS and pydoop are replaced "
           "by the local file system."),
    ("code", HEADER + "from distkeras_b200.data import Dataset\ntmp = tempfile.mkdtemp(prefix='dk_npy_')"),
    ("md", "## Structured `.npy` files"),
    ("code", "fields = np.dtype([('pt', np.float32), ('eta', np.float32), ('phi', np.float32), ('charge', np.int32)])\n"
             "rng = np.random.default_rng(0)\n"
             "for i in range(8):\n"
             "    rec = np.zeros(250, dtype=fields)\n"
             "    rec['pt'], rec['eta'] = rng.exponential(30.0, 250), rng.normal(0.0, 1.5, 250)\n"
             "    rec['phi'], rec['charge'] = rng.uniform(-np.pi, np.pi, 250), rng.choice([-1, 1], 250)\n"
             "    np.save(os.path.join(tmp, f'events_{i}.npy'), rec)\n"
             "file_paths = sorted(os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.npy'))\nprint(len(file_paths), 'files')"),
    ("md", "## A table of paths, repartitioned for parallelism"),
    ("code", "df = Dataset.from_rows([{'path_id': i} for i in range(len(file_paths))]).repartition(4)\n"
             "df.printSchema()\nprint(df.count(), 'paths in', df.rdd.getNumPartitions(), 'partitions')"),
    ("md", "## Prototype of the parser on one file"),
    ("code", "data = np.load(file_paths[0])\nnames = list(data.dtype.fields)\nprint(names, data[0])"),
    ("md", "## The partition mapper\nEvery partition opens its own files and emits one row per record."),
    ("code", "def parse(iterator):\n"
             "    for row in iterator:\n"
             "        rec = np.load(file_paths[int(row['path_id'])])\n"
             "        for r in rec:\n"
             "            yield {k: r[k].item() for k in names}\n"
             "dataset = df.rdd.mapPartitions(parse).toDF()\n"
             "dataset.printSchema()\nprint(dataset.count(), 'records;', dataset.take(1))"),
    ("md", "## Columnar fast path\nWhen every file has the same dtype the shards can be concatenated field by field without "
           "going through rows."),
    ("code", "cols = {k: np.concatenate([np.load(p)[k] for p in file_paths]) for k in names}\n"
             "fast = Dataset(cols, num_partitions=len(file_paths))\n"
             "assert fast.count() == dataset.count()\n"
             "features = np.stack([cols['pt'], cols['eta'], cols['phi']], axis=1)\n"
             "fast = fast.with_column('features', features)\n"
             "fast.to_parquet(os.path.join(tmp, 'events.parquet'))\nprint(fast)"),
]


# ------------------------------------------------------------------------------------------------------------------
NOTEBOOKS["kafka_spark_high_throughput_ml_pipeline"] = [
    ("md", "# High-throughput streaming inference\n\nCounterpart of the reference's `kafka_spark_high_throughput_ml_pipeline.ipynb`: a producer "
           "(`kafka_producer.py`) pushes JSON events to a topic; the pipeline collects them into micro-batches, turns the JSON into rows "
           "(`json_to_dataframe_row`), standardises the features with the statistics of the training set, scores them with "
           "`ModelPredictor`, indexes the prediction and keeps the events classified as signal.  There is no broker in this "
           "sandbox, so the topic is a spool directory (`--sink spool`); with the `kafka` package and a broker the same producer "
           "takes `--sink kafka`."),
    ("code", HEADER + "import glob, json, subprocess\n"
             "from distkeras_b200.data import Dataset, synthetic_higgs\n"
             "from distkeras_b200.models import Sequential, Dense, Dropout, Activation\n"
             "from distkeras_b200.trainers import SingleTrainer\n"
             "from distkeras_b200.predictors import ModelPredictor\n"
             "from distkeras_b200.transformers import LabelIndexTransformer, OneHotTransformer, StandardTransformer\n"
             "from distkeras_b200.utils import json_to_dataframe_row\n"
             "topic = tempfile.mkdtemp(prefix='Machine_Learning_')"),
    ("md", "## A model to serve\nThe 500-1000-500 perceptron of the reference notebook, trained for one pass on synthetic Higgs events."),
    ("code", "train = synthetic_higgs(20000)\n"
             "scaler = StandardTransformer(['features'])\n"
             "train = OneHotTransformer(2, 'label', 'label_encoded').transform(scaler.transform(train))\n"
             "mu, sd = scaler.means['features'].clone(), scaler.stddevs['features'].clone()   # statistics of the training set\n"
             "model = Sequential()\n"
             "model.add(Dense(500, input_shape=(30,)))\nmodel.add(Activation('relu'))\nmodel.add(Dropout(0.3))\n"
             "model.add(Dense(1000))\nmodel.add(Activation('relu'))\nmodel.add(Dropout(0.3))\n"
             "model.add(Dense(500))\nmodel.add(Activation('relu'))\n"
             "model.add(Dense(2))\nmodel.add(Activation('softmax'))\n"
             "trainer = SingleTrainer(model, 'adagrad', 'categorical_crossentropy', features_col='features_normalized',\n"
             "                        label_col='label_encoded', batch_size=64, num_epoch=1)\n"
             "model = trainer.train(train)\nprint('trained in %.2f s' % trainer.get_training_time())"),
    ("md", "## Start the producer\nA separate process, as in production: bursts of JSON messages every 0.2 s (5 s in the reference)."),
    ("code", "producer = subprocess.Popen([sys.executable, 'kafka_producer.py', '--sink', 'spool', '--dir', topic,\n"
             "                             '--bursts', '5', '--rows', '1000', '--interval', '0.2'])"),
    ("md", "## The streaming job\nEvery micro-batch interval: read what arrived, JSON -> rows -> `Dataset`, standardise, predict, index, filter."),
    ("code", "predictor = ModelPredictor(model, features_col='features_normalized')\n"
             "indexer = LabelIndexTransformer(output_dim=2)\n"
             "seen, total, signal = set(), 0, 0\n"
             "t0 = time.time()\n"
             "while True:\n"
             "    time.sleep(0.25)                                  # micro-batch interval (10 s in the reference)\n"
             "    files = sorted(f for f in glob.glob(os.path.join(topic, 'burst_*.jsonl')) if f not in seen)\n"
             "    rows = []\n"
             "    for f in files:\n"
             "        seen.add(f)\n"
             "        with open(f) as fh:\n"
             "            rows += [json_to_dataframe_row(line) for line in fh if line.strip()]\n"
             "    if rows:\n"
             "        batch = Dataset.from_rows(rows)\n"
             "        batch = batch.with_column('features_normalized', (batch['features'].float() - mu) / sd)\n"
             "        out = indexer.transform(predictor.predict(batch))\n"
             "        hits = out.filter(lambda d: d['prediction_index'] == 1.0).count()\n"
             "        total, signal = total + batch.count(), signal + hits\n"
             "        print('micro-batch: %5d events, %5d classified as signal' % (batch.count(), hits))\n"
             "    elif os.path.exists(os.path.join(topic, '_DONE')):\n"
             "        break\n"
             "producer.wait()\n"
             "print('%d events in %.2f s -> %.0f events/s, %d signal candidates' % (total, time.time() - t0, total / (time.time() - t0), signal))"),
]


def notebook(cells):
    out = []
    for kind, src in cells:
        lines = src.split("\n")
        source = [ln + "\n" for ln in lines[:-1]] + [lines[-1]]
        if kind == "md":
            out.append({"cell_type": "markdown", "metadata": {}, "source": source})
        else:
            out.append({"cell_type": "code", "metadata": {}, "execution_count": None, "outputs": [], "source": source})
    return {"cells": out, "metadata": {"kernelspec": {"display_name": "Python 3", "language": "python", "name": "python3"},
                                       "language_info": {"name": "python"}}, "nbformat": 4, "nbformat_minor": 5}


def main():
    for name, cells in NOTEBOOKS.items():
        path = os.path.join(ROOT, "examples", name + ".ipynb")
        with open(path, "w") as f:
            json.dump(notebook(cells), f, indent=1)
            f.write("\n")
        print("wrote", path)


if __name__ == "__main__":
    main()
