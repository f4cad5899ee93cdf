"""Packaging (reference: ``setup.py`` installs the pure-Python ``distkeras`` package).

``build_ext`` compiles the sm_100a kernels with ``build_native.py`` and places
``libdistkeras_b200.so`` inside the package (``distkeras_b200/lib/``), so an in-tree
``python setup.py build_ext --inplace`` / ``develop`` and a wheel both carry the native library.
"""
import os
import subprocess
import sys

from setuptools import Command, find_packages, setup
from setuptools.command.build_py import build_py

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNative(Command):
    description = "compile csrc/*.cu for sm_100a (nvcc) into distkeras_b200/lib/"
    user_options = [("inplace", "i", "ignored: the library is always built in-tree")]

    def initialize_options(self):
        self.inplace = 1

    def finalize_options(self):
        pass

    def run(self):
        subprocess.check_call([sys.executable, os.path.join(ROOT, "build_native.py")], cwd=ROOT)


class BuildPy(build_py):
    def run(self):
        self.run_command("build_ext")
        super().run()


setup(
    name="distkeras_b200",
    version="0.1.0",
    description="B200-native asynchronous parameter-server training (dist-keras capabilities on sm_100a)",
    packages=find_packages(include=["distkeras_b200", "distkeras_b200.*"]),
    package_data={"distkeras_b200": ["lib/*.so"]},
    python_requires=">=3.10",
    install_requires=["torch", "numpy"],
    scripts=["scripts/punchcard.py", "scripts/generate_secret.py"],
    cmdclass={"build_ext": BuildNative, "build_py": BuildPy},
)
