"""Recurrent layers (LSTM, ReLU / tanh RNN) of the B200 backend.

Mirrors the reference's `Rnn` object -- `CuDnn.Rnn` (Cuda/Source/Libs/CuDnnRnn.c:63-1100) as wrapped by
`CudaBackend.createRnn / acquireRnnParams / updateRnnParams` (Cuda/Backend.py:171-350) and driven by
`Backend/Dnn.py:299-333` and `Modules/RNN.py:122-166`: same constructor arguments, same `forward / backwardData /
backwardParams` methods, same parameter names (`wi wf wc wo ri rf rc ro` + `bw* br*` for the LSTM, `wi ri bwi bri` for
the plain RNNs) exposed as views into one flat weight blob `W`.

B200 design: the reference hands the whole sequence to cuDNN's legacy RNN API (not even buildable against cuDNN 9,
SURVEY F6).  Here a layer is GEMMs on the tcgen05 engine plus one fused pointwise kernel per time step:
  forward   G = X Wcat^T for ALL steps at once ((T*B x in) x (in x 4H));  per step  G[t] += h[t-1] Rcat^T  and the cell
            kernel (bias, gates, c[t], h[t]) in one pass over B x 4H
  backward  per step the cell kernel (dgates[t], dc) and  dh = dgates[t] Rcat;  then  dX = dG Wcat  for all steps at once
  params    dWcat = dG^T X,  dRcat = dG[1:]^T H[:-1]  (two large GEMMs),  db = column sums of dG
The blob packing is ours (SURVEY 7 "LSTM weight blob layout"): per layer the gate matrices stacked [i; f; c; o] so that the
four gates of a step are one GEMM; all matrices first, then all biases (bw, br per layer), like cuDNN.
Unidirectional, dropout-free layers only; GRU and bidirectional modes raise NotImplementedError.
"""
import numpy as np

from .driver import lib, check
from .gpuarray import GPUArray

_f32 = np.dtype(np.float32)

MODE_RELU, MODE_TANH, MODE_LSTM, MODE_GRU = 0, 1, 2, 3
DIR_UNI, DIR_BI = 0, 1

_GATES = {MODE_RELU: ("i", ), MODE_TANH: ("i", ), MODE_LSTM: ("i", "f", "c", "o")}


class RnnReserve:
	"""What the backward passes need from the forward pass (the reference's opaque `reserve` buffer)."""

	def __init__(self):
		self.layers = []          # per layer: dict(acts=(T,B,G*H) activations, cells=(T,B,H) or None, indata=(T,B,in))
		self.dgates = None        # filled by backwardData, consumed by backwardParams


class Rnn:
	def __init__(self, backend, insize, hsize, dtype, layers=1, algo=0, mode=MODE_LSTM, direction=DIR_UNI, dropout=0.0, seed=0,
				 batchsize=0):
		if np.dtype(dtype) != _f32:
			raise NotImplementedError("recurrent layers are float32 only (the reference creates them with np.float32, Dnn.py:301)")
		if mode == MODE_GRU:
			raise NotImplementedError("GRU mode is not implemented in the B200 backend yet")
		if direction != DIR_UNI:
			raise NotImplementedError("bidirectional recurrent layers are not implemented in the B200 backend yet")
		if mode not in _GATES:
			raise ValueError("invalid rnn mode %s" % mode)
		if dropout != 0.0 and layers > 1:
			raise NotImplementedError("dropout between recurrent layers is not implemented in the B200 backend yet")

		self.backend = backend
		self.insize, self.hsize, self.layers = int(insize), int(hsize), int(layers)
		self.dtype, self.algo, self.mode, self.direction = _f32, algo, mode, direction
		self.dropout, self.seed, self.batchsize = dropout, seed, batchsize
		self.ngates = len(_GATES[mode])

		# blob layout: [layer: Wcat (G*H, in_l) | Rcat (G*H, H)] ... [layer: bw (G*H) | br (G*H)] ...
		H, G = self.hsize, self.ngates
		self.matOffsets, self.biasOffsets = [], []
		off = 0
		for layer in range(self.layers):
			insz = self.insize if layer == 0 else H
			self.matOffsets.append((off, off + G * H * insz))
			off += G * H * (insz + H)
		for layer in range(self.layers):
			self.biasOffsets.append((off, off + G * H))
			off += 2 * G * H
		self.wsize = off

	# ------------------------------------------------------------------------------------------ parameter views
	def _view(self, W, offset, shape):
		size = int(np.prod(shape))
		itemsize = W.dtype.itemsize
		return GPUArray(shape, W.dtype, gpudata=W.gpudata[offset * itemsize:(offset + size) * itemsize])

	def layerInsize(self, layer):
		return self.insize if layer == 0 else self.hsize

	def stacked(self, W, layer):
		"""(Wcat, Rcat, bw, br) of one layer: the stacked gate matrices the GEMMs use."""
		H, G, insz = self.hsize, self.ngates, self.layerInsize(layer)
		woff, roff = self.matOffsets[layer]
		bwoff, broff = self.biasOffsets[layer]
		return (self._view(W, woff, (G * H, insz)), self._view(W, roff, (G * H, H)), self._view(W, bwoff, (G * H, )),
				self._view(W, broff, (G * H, )))

	def getParam(self, W, layer, linLayer):
		"""((Woffset, wsize), (biasOffset, biasSize)) in elements, as CuDnn.Rnn.getParam returns (Cuda/Backend.py:205-218)."""
		H, G, insz = self.hsize, self.ngates, self.layerInsize(layer)
		woff, roff = self.matOffsets[layer]
		bwoff, broff = self.biasOffsets[layer]
		if linLayer < G:
			return (woff + linLayer * H * insz, H * insz), (bwoff + linLayer * H, H)
		g = linLayer - G
		return (roff + g * H * H, H * H), (broff + g * H, H)

	def acquireParams(self, W):
		"""List (one dict per layer) of named views into W -- the reference's acquireRnnParams."""
		H, G = self.hsize, self.ngates
		params = []
		for layer in range(self.layers):
			insz = self.layerInsize(layer)
			layerparams = {}
			for linLayer in range(2 * G):
				wtype = "w" if linLayer < G else "r"
				gate = _GATES[self.mode][linLayer % G]
				(woff, wsize), (boff, bsize) = self.getParam(W, layer, linLayer)
				layerparams["%s%s" % (wtype, gate)] = self._view(W, woff, (H, insz if wtype == "w" else H))
				layerparams["b%s%s" % (wtype, gate)] = self._view(W, boff, (bsize, ))
			params.append(layerparams)
		return params

	# ------------------------------------------------------------------------------------------ forward
	def forward(self, data, W, hidden=None, cells=None, test=False, allocator=None):
		if data.ndim != 3 or data.shape[2] != self.insize or data.dtype != _f32:
			raise ValueError("invalid rnn input layout %s" % (data.shape, ))
		if W.dtype != _f32 or W.size != self.wsize:
			raise ValueError("invalid rnn weights size")
		T, B, _ = data.shape
		H, G = self.hsize, self.ngates
		blas = self.backend.blas
		reserve = RnnReserve()

		x = data
		for layer in range(self.layers):
			insz = self.layerInsize(layer)
			Wcat, Rcat, bw, br = self.stacked(W, layer)
			h0 = None if hidden is None else hidden[layer]
			c0 = None if cells is None else cells[layer]

			# input projection of every step in one GEMM: (T*B, in) x (in, G*H)
			gates = GPUArray((T, B, G * H), _f32, allocator=allocator)
			blas.gemm(x.reshape(T * B, insz), Wcat, gates.reshape(T * B, G * H), transpB=True)

			# LSTM: separate output buffer; plain RNN: the pre-activation buffer becomes the output in place
			out = GPUArray((T, B, H), _f32, allocator=allocator) if self.mode == MODE_LSTM else gates
			cellsbuf = GPUArray((T, B, H), _f32, allocator=allocator) if self.mode == MODE_LSTM else None

			for t in range(T):
				hprev = h0 if t == 0 else out[t - 1]
				if hprev is not None:
					blas.gemm(hprev, Rcat, gates[t], transpB=True, alpha=1.0, beta=1.0)
				if self.mode == MODE_LSTM:
					cprev = c0 if t == 0 else cellsbuf[t - 1]
					check(lib.pz_lstm_cell_fwd(gates[t].ptr, bw.ptr, br.ptr, cprev.ptr if cprev is not None else None, cellsbuf[t].ptr,
											   out[t].ptr, B, H, None))
				else:
					check(lib.pz_rnn_cell_fwd(gates[t].ptr, bw.ptr, br.ptr, B, H, self.mode, None))

			reserve.layers.append({"acts": gates, "cells": cellsbuf, "indata": x, "out": out, "h0": h0, "c0": c0})
			x = out

		return x if test else (x, reserve)

	# ------------------------------------------------------------------------------------------ backward (data)
	def backwardData(self, grad, outdata, W, reserve, hidden=None, cells=None, allocator=None):
		T, B, H = outdata.shape
		G = self.ngates
		if grad.shape != outdata.shape:
			raise ValueError("invalid rnn gradient layout %s" % (grad.shape, ))
		blas = self.backend.blas
		reserve.dgates = [None] * self.layers
		dhx, dcx = [None] * self.layers, [None] * self.layers

		dy = grad
		for layer in range(self.layers - 1, -1, -1):
			saved = reserve.layers[layer]
			insz = self.layerInsize(layer)
			Wcat, Rcat, _, _ = self.stacked(W, layer)
			acts, cellsbuf, out, h0, c0 = saved["acts"], saved["cells"], saved["out"], saved["h0"], saved["c0"]

			dgates = GPUArray((T, B, G * H), _f32, allocator=allocator)
			dhnext = GPUArray((B, H), _f32, allocator=allocator)
			dc = GPUArray((B, H), _f32, allocator=allocator) if self.mode == MODE_LSTM else None

			for t in range(T - 1, -1, -1):
				last = t == T - 1
				if self.mode == MODE_LSTM:
					cprev = c0 if t == 0 else cellsbuf[t - 1]
					check(lib.pz_lstm_cell_bwd(dy[t].ptr, None if last else dhnext.ptr, dc.ptr, acts[t].ptr, cellsbuf[t].ptr,
											   cprev.ptr if cprev is not None else None, dgates[t].ptr, B, H, 1 if last else 0, None))
				else:
					check(lib.pz_rnn_cell_bwd(dy[t].ptr, None if last else dhnext.ptr, out[t].ptr, dgates[t].ptr, B * H, self.mode, None))
				# gradient reaching h[t-1] through the recurrence: (B, G*H) x (G*H, H)
				blas.gemm(dgates[t], Rcat, dhnext)

			dhx[layer], dcx[layer] = dhnext, dc
			reserve.dgates[layer] = dgates

			# input gradient of every step in one GEMM: (T*B, G*H) x (G*H, in)
			dx = GPUArray((T, B, insz), _f32, allocator=allocator)
			blas.gemm(dgates.reshape(T * B, G * H), Wcat, dx.reshape(T * B, insz))
			dy = dx

		return dy, dhx, dcx

	# ------------------------------------------------------------------------------------------ backward (params)
	def backwardParams(self, data, outdata, reserve, hidden=None, allocator=None):
		if reserve.dgates is None:
			raise ValueError("backwardParams needs the reserve of a backwardData call")
		T, B, H = outdata.shape
		G = self.ngates
		blas, matmod = self.backend.blas, self.backend.matmod
		dw = GPUArray.zeros((self.wsize, ), _f32, allocator=allocator)

		for layer in range(self.layers):
			saved = reserve.layers[layer]
			insz = self.layerInsize(layer)
			dWcat, dRcat, dbw, dbr = self.stacked(dw, layer)
			dgates, x, out, h0 = reserve.dgates[layer], saved["indata"], saved["out"], saved["h0"]

			dg2 = dgates.reshape(T * B, G * H)
			blas.gemm(dg2, x.reshape(T * B, insz), dWcat, transpA=True)                                   # (G*H, in)
			if T > 1:
				blas.gemm(dgates[1:].reshape((T - 1) * B, G * H), out[:T - 1].reshape((T - 1) * B, H), dRcat, transpA=True)
			if h0 is not None:
				blas.gemm(dgates[0], h0, dRcat, transpA=True, alpha=1.0, beta=1.0 if T > 1 else 0.0)
			matmod.matsum(dg2, axis=0, out=dbw)
			dbr.set(dbw)

		return dw
