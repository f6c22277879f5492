"""Import the REAL reference (read-only tree at /root/reference, or its staged copy baseline/_ref on a GPU box),
with stub modules for the dependencies this image lacks.

TEST / BENCH-BASELINE INFRASTRUCTURE ONLY, usable where /root/reference or baseline/_ref exists:
it is used by oracle/make_golden.py to mint tests/golden/ and by the `reference`-marked tests
that pin oracle/oracle.py against the reference itself.  Nothing is copied from the reference.
"""
import importlib
import importlib.machinery
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_DIR = os.path.join(ROOT, 'baseline', '_ref')  # git-ignored copy made by stage(); travels to the GPU box with the snapshot


def _reference_dir():
	for d in (os.environ.get('CONVASR_REFERENCE_DIR'), '/root/reference', STAGED_DIR):
		if d and os.path.isfile(os.path.join(d, 'models.py')):
			return d
	return '/root/reference'


REFERENCE_DIR = _reference_dir()


def stage():
	"""Copy the reference's top-level Python modules, UNMODIFIED, into the git-ignored baseline/_ref/ so that bench.py's
	reference arm (and the gpu_baseline leg) can run the real reference on the GPU box, where /root/reference does not
	exist.  Build-container only; nothing under baseline/_ref is tracked."""
	import shutil
	src = '/root/reference'
	if not os.path.isfile(os.path.join(src, 'models.py')):
		return None
	os.makedirs(STAGED_DIR, exist_ok = True)
	for f in sorted(os.listdir(src)):
		if f.endswith('.py'):
			shutil.copyfile(os.path.join(src, f), os.path.join(STAGED_DIR, f))
	return STAGED_DIR


def available():
	return os.path.isfile(os.path.join(REFERENCE_DIR, 'models.py'))


_CACHE = {}


def load():
	"""Returns a namespace with the reference modules: models, ctc, decoders, transcript_generators,
	text_tokenizers."""
	if _CACHE:
		return types.SimpleNamespace(**_CACHE)
	if not available():
		raise RuntimeError(f'reference tree not found at {REFERENCE_DIR}')
	import numpy as np
	import torchaudio

	for name in ('onnxruntime', 'apex', 'librosa', 'soundfile', 'Levenshtein'):
		if name not in sys.modules:
			m = types.ModuleType(name)
			m.__spec__ = importlib.machinery.ModuleSpec(name, None)  # torch._dynamo needs a real spec
			sys.modules[name] = m
	librosa = sys.modules['librosa']
	# the only missing arithmetic: librosa's Slaney mel basis == torchaudio's (6.5e-8)
	librosa.filters = types.SimpleNamespace(
		mel = lambda sr, n_fft, n_mels = 128, fmin = 0., fmax = None: torchaudio.functional.
		melscale_fbanks(n_fft // 2 + 1, float(fmin), float(fmax), n_mels, sr, norm = 'slaney', mel_scale = 'slaney').T.numpy()
	)
	librosa.util = types.SimpleNamespace(pad_center = lambda w, size: np.pad(w, ((size - len(w)) // 2, ) * 2))

	# the reference modules are imported by bare name; keep them out of the way of the drop-in
	# modules of the same names by importing under a private prefix
	saved_path = list(sys.path)
	saved_modules = {n: sys.modules.get(n) for n in ('models', 'ctc', 'decoders', 'transcript_generators', 'text_tokenizers', 'shaping', 'transcripts', 'datasets', 'optimizers', 'audio', 'utils', 'text_processing')}
	for n in saved_modules:
		sys.modules.pop(n, None)
	sys.path.insert(0, REFERENCE_DIR)
	try:
		for n in ('shaping', 'transcripts', 'models', 'ctc', 'decoders', 'transcript_generators', 'text_tokenizers', 'optimizers'):
			_CACHE[n] = importlib.import_module(n)
		try:  # the batch feed's collate_fn (datasets.py:305-332); pulls in text_processing / audio / utils
			import warnings
			with warnings.catch_warnings():
				warnings.simplefilter('ignore')
				_CACHE['datasets'] = importlib.import_module('datasets')
		except Exception:
			_CACHE['datasets'] = None
	finally:
		sys.path[:] = saved_path
		for n, m in saved_modules.items():
			sys.modules.pop(n, None)
			if m is not None:
				sys.modules[n] = m
	return types.SimpleNamespace(**_CACHE)


def make_model(name = 'Wav2Letter', num_classes = (38, ), frontend = True, **kwargs):
	ref = load()
	fe = ref.models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window') if frontend else None
	model = getattr(ref.models, name)(64, list(num_classes), frontend = fe, dropout = 0., **kwargs)
	model.check_time_dim_padded = False  # Wav2Letter's ctor does not accept the kwarg (models.py:820-835)
	return model.eval()
