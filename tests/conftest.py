import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
	config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu on the GPU box')
	config.addinivalue_line('markers', 'reference: needs the read-only reference tree at /root/reference (build container only)')


def pytest_collection_modifyitems(config, items):
	import torch
	from oracle import reference_shim
	has_gpu = torch.cuda.is_available()
	has_ref = reference_shim.available()
	for item in items:
		if 'gpu' in item.keywords and not has_gpu:
			item.add_marker(pytest.mark.skip(reason = 'no CUDA device'))
		if 'reference' in item.keywords and not has_ref:
			item.add_marker(pytest.mark.skip(reason = 'reference tree not present'))


@pytest.fixture(scope = 'session')
def golden():
	import torch

	def load(name):
		return torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only = False)

	return load
