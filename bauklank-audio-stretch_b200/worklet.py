"""Host-side mirror of the reference's worklet control plane.

``WorkletTimeline`` restates ``WasmProcessor`` minus the DSP: the ``remoteMethods`` RPC surface
(app/SignalsmithStretch.mjs:603-744 -- ``configure latency setUpdateInterval stop start schedule dropBuffers
addBuffers``), the time map they edit, and the per-quantum bookkeeping of ``process()`` (:840-849, :883-897).  It is
pure arithmetic on JS numbers (Python floats are the same IEEE doubles), so a recorded control trace -- "at render
quantum k, call schedule({...})" -- resolves into the per-quantum table the batched engine consumes
(``bsb_add_kiosk_table``), or drives any engine with the reference's 18-call surface quantum by quantum exactly as the
worklet would (``render``), which is how the tests tie the two together.

``ControllerMapper`` restates how the kiosk app turns the hardware controller's messages
(``{"type":"set","channel":"A","key":"rate|volume|tone",...}``, server-multi.py:47-48, normalised at :722-737) into
``schedule()`` calls (app/multi/app.mjs:537-616 ``applyIncomingSet``, :478-507 ``controlsChanged``).

Reference quirks kept on purpose:
* ``schedule`` inherits only active/rate/semitones/loopStart/loopEnd from the latest segment (:670-678); a call that
  omits tonalityHz / formantSemitones / formantCompensation / formantBaseHz would hand NaN to the engine.  The kiosk
  app always passes them (app/multi/app.mjs:495-507); this mirror raises instead of propagating NaN.
* the buffer-fill loop (:900-931) only works for a single stored buffer (with two, ``count`` overruns the view and
  ``Float32Array.set`` throws); one ``addBuffers`` call is supported, like the kiosk makes (app/multi/app.mjs:369-376).
"""
import ctypes as C
import json
import math

import numpy as np

from . import _capi

QUANTUM = 128  # render quantum of the Web Audio API


def _js_round(x):
    return int(math.floor(x + 0.5))


def _clamp(x, lo, hi):
    return max(lo, min(hi, x))


def _finite(v, default):
    """toFiniteNumber of the app: Number(v) if finite else default."""
    try:
        n = float(v)
    except (TypeError, ValueError):
        return default
    return n if math.isfinite(n) else default


class WorkletTimeline:
    FIELDS = ("active", "input", "output", "rate", "semitones", "tonalityHz", "formantSemitones", "formantCompensation",
              "formantBaseHz", "loopStart", "loopEnd")

    def __init__(self, sample_rate=48000.0, channels=2, config=None, lib=None):
        self.sample_rate = float(sample_rate)
        self.channels = int(channels)
        self.lib = lib
        self.current_time = 0.0                     # AudioWorkletGlobalScope.currentTime
        self.quantum_index = 0
        self.audio = None                           # the single stored buffer [channels, n] (or None)
        self.audio_start = 0                        # audioBuffersStart
        self.audio_end = 0                          # audioBuffersEnd
        self.time_interval_samples = self.sample_rate * 0.1
        self.time_map = [dict(active=False, input=0.0, output=0.0, rate=1.0, semitones=0.0, tonalityHz=8000.0,
                              formantSemitones=0.0, formantCompensation=False, formantBaseHz=0.0, loopStart=0.0,
                              loopEnd=0.0)]           # :587-600
        self.config = dict(preset="default")        # :782-784
        if config:
            self.config.update(config)
        self._geometry()

    # ---- configure (:786-801): only the latencies matter here
    def _geometry(self):
        sr = self.sample_rate
        if self.config.get("blockMs"):
            block = _js_round(self.config["blockMs"] / 1000 * sr)
            interval = _js_round((self.config.get("intervalMs") or self.config["blockMs"] * 0.25) / 1000 * sr)
            split = 1 if self.config.get("splitComputation") else 0
        elif self.config.get("preset") == "cheaper":
            d = float(np.float32(sr)); block, interval, split = int(d * 0.1), int(d * 0.04), 1
        else:
            d = float(np.float32(sr)); block, interval, split = int(d * 0.12), int(d * 0.03), 0
        lib = self.lib or _capi.load_library()
        out = (C.c_int * 6)()
        if lib.bsb_query_geometry(block, interval, split, out) != 0:
            raise ValueError("unsupported block/interval")
        self.block_samples, self.interval_samples, self.split = block, interval, split
        self.input_latency, self.output_latency = int(out[2]), int(out[3])
        self.input_latency_seconds = self.input_latency / sr
        self.output_latency_seconds = self.output_latency / sr
        self.buffer_length = self.input_latency + self.output_latency

    # ---- remoteMethods
    def configure(self, config):
        self.config.update(config)
        self._geometry()

    def latency(self):
        return self.input_latency_seconds + self.output_latency_seconds

    def setUpdateInterval(self, seconds):
        self.time_interval_samples = self.sample_rate * seconds

    def stop(self, when=None):
        if not isinstance(when, (int, float)):
            when = self.current_time
        return self.schedule(dict(active=False, output=when))

    def start(self, when=None, offset=None, duration=None, rate=None, semitones=None):
        if isinstance(when, dict):
            if "active" not in when:
                when["active"] = True
            return self.schedule(when)
        obj = dict(active=True, input=0.0, output=self.current_time + self.output_latency_seconds)
        if isinstance(when, (int, float)):
            obj["output"] = when
        if isinstance(offset, (int, float)):
            obj["input"] = offset
        if isinstance(rate, (int, float)):
            obj["rate"] = rate
        if isinstance(semitones, (int, float)):
            obj["semitones"] = semitones
        result = self.schedule(obj)
        if isinstance(duration, (int, float)):
            self.stop(obj["output"] + duration)
            obj["output"] += duration
            obj["active"] = False
            self.schedule(obj)
        return result

    def schedule(self, obj_in, adjust_previous=False):
        """:656-701, statement for statement."""
        tm = self.time_map
        output_time = obj_in["outputTime"] if "outputTime" in obj_in else self.current_time
        latest = tm[-1]
        while tm and tm[-1]["output"] >= output_time:
            latest = tm.pop()
        obj = dict(active=latest["active"], input=None, output=output_time, rate=latest["rate"],
                   semitones=latest["semitones"], loopStart=latest["loopStart"], loopEnd=latest["loopEnd"])
        obj.update(obj_in)
        if obj["input"] is None:
            rate = latest["rate"] if latest["active"] else 0
            obj["input"] = latest["input"] + (obj["output"] - latest["output"]) * rate
        tm.append(obj)
        if adjust_previous and len(tm) > 1:
            prev = tm[-2]
            if prev["output"] < self.current_time:
                rate = prev["rate"] if prev["active"] else 0
                prev["input"] += (self.current_time - prev["output"]) * rate
                prev["output"] = self.current_time
            prev["rate"] = (obj["input"] - prev["input"]) / (obj["output"] - prev["output"])
        while len(tm) > 1 and tm[1]["output"] <= output_time:
            tm.pop(0)
        return obj

    def addBuffers(self, sample_buffers):
        buf = np.ascontiguousarray(np.asarray(sample_buffers, np.float32))
        if buf.ndim == 1:
            buf = buf[None, :]
        if self.audio is not None:
            raise NotImplementedError("one stored buffer only: the reference's fill loop (:900-931) overruns its view with two")
        self.audio = buf
        self.audio_end += buf.shape[1]
        return self.audio_end / self.sample_rate

    def dropBuffers(self, to_seconds=None):
        if not isinstance(to_seconds, (int, float)):
            self.audio = None
            self.audio_start = self.audio_end = 0
            return dict(start=0, end=0)
        if self.audio is not None and (self.audio_start + self.audio.shape[1]) / self.sample_rate <= to_seconds:
            self.audio_start += self.audio.shape[1]
            self.audio = None
        return dict(start=self.audio_start / self.sample_rate, end=self.audio_end / self.sample_rate)

    # ---- process(), control part (:840-849, :883-897)
    def quantum(self):
        """Bookkeeping of one render quantum; returns the resolved record and advances currentTime."""
        sr = self.sample_rate
        output_time = self.current_time + self.output_latency_seconds
        tm = self.time_map
        while len(tm) > 1 and tm[1]["output"] <= output_time:
            tm.pop(0)
        seg = tm[0]
        for k in ("tonalityHz", "formantSemitones", "formantCompensation", "formantBaseHz"):
            if k not in seg:
                raise ValueError("time-map segment without %s: the reference would pass NaN to the engine (schedule() does "
                                 "not inherit it, app/SignalsmithStretch.mjs:670-678)" % k)
        rec = dict(active=bool(seg["active"]), rate=float(seg["rate"]),
                   semitones=np.float32(seg["semitones"]), tonality_limit=np.float32(seg["tonalityHz"] / sr),
                   formant_semitones=np.float32(seg["formantSemitones"]), formant_compensation=bool(seg["formantCompensation"]),
                   formant_base=np.float32(seg["formantBaseHz"] / sr), input_samples_end=0, input_time=None,
                   valid_start=self.audio_start if self.audio is not None else 0,
                   valid_end=(self.audio_start + self.audio.shape[1]) if self.audio is not None else 0)
        if seg["active"]:
            input_time = seg["input"] + (output_time - seg["output"]) * seg["rate"]
            loop_length = seg["loopEnd"] - seg["loopStart"]
            if loop_length > 0 and input_time >= seg["loopEnd"]:
                seg["input"] -= loop_length
                input_time -= loop_length
            input_time += self.input_latency_seconds
            rec["input_time"] = input_time
            rec["input_samples_end"] = _js_round(input_time * sr)
        self.current_time = (self.quantum_index + 1) * QUANTUM / sr   # currentTime = frames rendered / sampleRate
        self.quantum_index += 1
        return rec

    def _apply_events(self, events, ei):
        while ei < len(events) and events[ei][0] <= self.quantum_index:
            _, method, args = events[ei]
            getattr(self, method)(*args)
            ei += 1
        return ei

    def resolve(self, n_out, events=()):
        """Run the control plane for ceil(n_out/128) quanta; ``events`` = sorted [(quantum_index, method, args)], each
        applied before that quantum renders (messages are handled between render calls).  Returns the records."""
        events = sorted(events, key=lambda e: e[0])
        ei, recs = 0, []
        for _ in range((n_out + QUANTUM - 1) // QUANTUM):
            ei = self._apply_events(events, ei)
            recs.append(self.quantum())
        return recs

    @staticmethod
    def table(recs):
        """ctypes array of ``bsb_quantum`` for ``bsb_add_kiosk_table``."""
        arr = (_capi.Quantum * len(recs))()
        for q, r in zip(arr, recs):
            q.rate, q.input_samples_end = r["rate"], r["input_samples_end"]
            q.valid_start, q.valid_end = r["valid_start"], r["valid_end"]
            q.semitones, q.tonality_limit = r["semitones"], r["tonality_limit"]
            q.formant_semitones, q.formant_base = r["formant_semitones"], r["formant_base"]
            q.formant_compensation, q.active = int(r["formant_compensation"]), int(r["active"])
            q.transpose_factor = q.formant_factor = float("nan")
        return arr

    def render(self, engine, n_out, events=(), clip=None):
        """Drive an engine with the reference's 18-call surface exactly like ``process()`` does (buffer playback and
        inactive branches, :861-869 and :883-943).  ``clip``: the stored audio if it is not added through events."""
        sr = self.sample_rate
        ch = self.channels
        if self.config.get("blockMs"):
            engine.configure(ch, self.block_samples, self.interval_samples, self.split); engine.reset()
        elif self.config.get("preset") == "cheaper":
            engine.presetCheaper(ch, sr)
        else:
            engine.presetDefault(ch, sr)
        engine.setBuffers(ch, self.buffer_length)
        if clip is not None:
            self.addBuffers(clip)
        events = sorted(events, key=lambda e: e[0])
        out = np.zeros((ch, n_out), np.float32)
        ei = pos = 0
        while pos < n_out:
            ei = self._apply_events(events, ei)
            r = self.quantum()
            q = min(QUANTUM, n_out - pos)
            engine.setTransposeSemitones(float(r["semitones"]), float(r["tonality_limit"]))
            engine.setFormantSemitones(float(r["formant_semitones"]), r["formant_compensation"])
            engine.setFormantBase(float(r["formant_base"]))
            ins, _ = engine.io_views()
            if not r["active"]:
                ins[:, :q] = 0
                engine.process(q, q)
            else:
                end = r["input_samples_end"]
                n = self.buffer_length
                ins[:] = 0
                if self.audio is not None:
                    lo, hi = max(end - n, r["valid_start"]), min(end, r["valid_end"])
                    if hi > lo:
                        ins[:, lo - (end - n):hi - (end - n)] = self.audio[:, lo - self.audio_start:hi - self.audio_start]
                engine.seek(n, r["rate"])
                engine.process(0, q)
            _, outs = engine.io_views()
            out[:, pos:pos + q] = outs[:, :q]
            pos += q
        return out


class ControllerMapper:
    """One kiosk engine's control state and how controller messages edit it (app/multi/app.mjs)."""

    CONTROL_DEFAULTS = dict(volume=0.10, active=True, rate=0.001, semitones=0, tonalityHz=16000, formantSemitones=0,
                            formantCompensation=False, formantBaseHz=200, loopStart=1, loopEnd=1)   # :106-123

    def __init__(self, audio_duration, channel="A"):
        self.values = dict(self.CONTROL_DEFAULTS)
        self.values["pan"] = -1 if channel == "A" else (1 if channel == "B" else 0)
        self.channel = channel
        self.audio_duration = float(audio_duration)

    @staticmethod
    def normalize(msg):
        """server-multi.py:722-737 ``_normalize_set_value``."""
        msg = dict(msg)
        key = str(msg.get("key", ""))
        if "value" in msg:
            try:
                if key in ("volume", "tone"):
                    msg["value"] = int(msg["value"])
                elif key == "rate":
                    msg["value"] = float(msg["value"])
            except (TypeError, ValueError):
                pass
        return msg

    def schedule_args(self, current_time, schedule_ahead=True):
        """controlsChanged (:478-507): the object handed to ``stretch.schedule``."""
        v = self.values
        d = self.audio_duration
        return dict(active=bool(v["active"]),
                    rate=_clamp(_finite(v["rate"], 0.001), 0.00001, 2),
                    semitones=_clamp(_finite(v["semitones"], 0), -48, 48),
                    tonalityHz=_clamp(_finite(v["tonalityHz"], 16000), 20, 22050),
                    formantSemitones=_clamp(_finite(v["formantSemitones"], 0), -48, 48),
                    formantCompensation=bool(v["formantCompensation"]),
                    formantBaseHz=_clamp(_finite(v["formantBaseHz"], 200), 20, 2000),
                    loopStart=_clamp(_finite(v["loopStart"], 0), 0, d),
                    loopEnd=_clamp(_finite(v["loopEnd"], d), 0, d),
                    outputTime=current_time + (0.1 if schedule_ahead else 0.0))

    def apply_set(self, key, value):
        """applyIncomingSet (:537-616).  Returns True if a schedule() call follows (controlsChanged)."""
        v = self.values
        if key == "volume":
            n = _finite(value, float("nan"))
            if not math.isfinite(n):
                return False
            v["volume"] = _clamp(n, 0, 100) / 100
            return True
        if key == "volumePercent":
            v["volume"] = _clamp(_finite(value, 100), 1, 100) / 100
            return True
        if key == "pan":
            n = _finite(value, float("nan"))
            if not math.isfinite(n):
                return False
            v["pan"] = _clamp((n / 50) - 1 if 0 <= n <= 100 else n, -1, 1)
            return True
        if key in ("tone", "semitones"):
            n = _finite(value, float("nan"))
            if not math.isfinite(n):
                return False
            lim = 24 if key == "tone" else 48
            v["semitones"] = _js_round(_clamp(n, -lim, lim))      # Math.round
            return True
        if key in v:
            cur = v[key]
            if isinstance(cur, bool):
                v[key] = bool(value)
            elif isinstance(cur, (int, float)):
                n = _finite(value, float("nan"))
                if not math.isfinite(n):
                    return False
                v[key] = n
            else:
                v[key] = value
            return True
        return False

    def trace_to_events(self, lines, sample_rate=48000.0):
        """Serial lines ``(t_seconds, json_text)`` -> sorted events for ``WorkletTimeline.resolve`` / ``render``: each
        accepted ``set`` for this channel becomes one ``schedule`` call made at the first render quantum whose
        currentTime is >= t (the message is handled between render calls)."""
        msgs = []
        for t, text in lines:
            try:
                msgs.append((t, json.loads(text)))
            except (ValueError, TypeError):
                continue
        return self.messages_to_events(msgs, sample_rate)

    def messages_to_events(self, msgs, sample_rate=48000.0):
        """The same for already parsed messages ``(t_seconds, dict)`` (what the bridge hands on after json.loads)."""
        events = []
        for t, msg in msgs:
            if not isinstance(msg, dict):
                continue
            msg = self.normalize(msg)
            if msg.get("type") != "set" or msg.get("channel", self.channel) != self.channel:
                continue
            if not self.apply_set(str(msg.get("key", "")), msg.get("value")):
                continue
            k = int(math.ceil(t * sample_rate / QUANTUM))
            events.append((k, "schedule", (self.schedule_args(k * QUANTUM / sample_rate),)))
        return events
