// Golden-vector dumper for libjtdsp parity (SURVEY 8f-1): drop this file into internal/processor/ of a jivetalking checkout
// with Go 1.26 + libffmpeg.a (the one thing this repository's build image does not have) and run
//
//	JT_GOLDEN_DIR=/tmp/jt_golden go test -run TestDumpJivetalkingGoldens ./internal/processor
//
// For each case it keeps the input WAV and writes <case>.json: one record per frame pulled from the buffersink with
// nb_samples and EVERY metadata key the real FFmpeg filters attached (decimal strings, untouched), the PCM the sink
// produced (<case>.pcm, raw, in the sink's own sample format) and, where the spec holds loudnorm, nothing more is needed:
// its JSON goes to the stats_file named in the spec.  scripts/compare_reference_golden.py (in the libjtdsp repository)
// replays the same WAV and spec strings through libjtdsp and compares with the tolerances of its own parity tests,
// turning "restated-oracle parity" into parity with the reference itself.
//
// This file is NOT compiled anywhere in the libjtdsp repository (no Go toolchain there); it only uses identifiers the
// reference already has (generateTestAudio, setupFilterGraph, runFilterGraph, FrameLoopConfig, DefaultFilterConfig,
// deriveEffectiveFilterConfig, Pass1FilterOrder, audio.OpenAudioFile, ffmpeg.AVDictIterate).
package processor

import (
	"context"
	"encoding/json"
	"os"
	"path/filepath"
	"testing"
	"unsafe"

	ffmpeg "github.com/linuxmatters/ffmpeg-statigo"
	"github.com/linuxmatters/jivetalking/internal/audio"
)

type goldenFrame struct {
	NbSamples int               `json:"nb_samples"`
	Format    int               `json:"format"`
	Meta      map[string]string `json:"meta"`
}

type goldenCase struct {
	Name   string        `json:"name"`
	Spec   string        `json:"spec"`
	WAV    string        `json:"wav"`
	Frames []goldenFrame `json:"frames"`
	PCM    string        `json:"pcm,omitempty"`
}

func dumpGolden(t *testing.T, dir, name, wav, spec string, keepPCM bool) {
	t.Helper()
	reader, _, err := audio.OpenAudioFile(wav)
	if err != nil {
		t.Fatal(err)
	}
	defer reader.Close()
	graph, src, sink, err := setupFilterGraph(reader.DecoderContext(), spec)
	if err != nil {
		t.Fatal(err)
	}
	defer ffmpeg.AVFilterGraphFree(&graph)
	gc := goldenCase{Name: name, Spec: spec, WAV: filepath.Base(wav)}
	var pcm *os.File
	if keepPCM {
		gc.PCM = name + ".pcm"
		if pcm, err = os.Create(filepath.Join(dir, gc.PCM)); err != nil {
			t.Fatal(err)
		}
		defer pcm.Close()
	}
	bytesPer := map[int]int{1: 2, 2: 4, 3: 4, 4: 8} // AV_SAMPLE_FMT_S16 / S32 / FLT / DBL, mono sinks only
	err = runFilterGraph(context.Background(), reader, src, sink, FrameLoopConfig{
		OnFrame: func(_, f *ffmpeg.AVFrame) error {
			rec := goldenFrame{NbSamples: f.NbSamples(), Format: f.Format(), Meta: map[string]string{}}
			var e *ffmpeg.AVDictionaryEntry
			for {
				if e = ffmpeg.AVDictIterate(f.Metadata(), e); e == nil {
					break
				}
				rec.Meta[e.Key().String()] = e.Value().String()
			}
			gc.Frames = append(gc.Frames, rec)
			if pcm != nil {
				n := f.NbSamples() * bytesPer[f.Format()]
				if _, err := pcm.Write(unsafe.Slice((*byte)(f.Data().Get(0)), n)); err != nil {
					return err
				}
			}
			return nil
		},
	})
	if err != nil {
		t.Fatal(err)
	}
	out, _ := json.Marshal(gc)
	if err := os.WriteFile(filepath.Join(dir, name+".json"), out, 0o644); err != nil {
		t.Fatal(err)
	}
}

func TestDumpJivetalkingGoldens(t *testing.T) {
	dir := os.Getenv("JT_GOLDEN_DIR")
	if dir == "" {
		t.Skip("set JT_GOLDEN_DIR to dump golden vectors")
	}
	if err := os.MkdirAll(dir, 0o755); err != nil {
		t.Fatal(err)
	}
	// the recipes of analyser_test.go:135-148 and benchmark_test.go:96-109, at both rates
	inputs := map[string]TestAudioOptions{
		"tone48": {SampleRate: 48000, DurationSecs: 20, ToneFreq: 440, ToneLevel: -23, NoiseLevel: -60},
		"tone44": {SampleRate: 44100, DurationSecs: 20, ToneFreq: 440, ToneLevel: -23, NoiseLevel: -60},
	}
	pass1 := deriveEffectiveFilterConfig(DefaultFilterConfig())
	pass1.FilterOrder = cloneFilterOrder(Pass1FilterOrder)
	pass2 := deriveEffectiveFilterConfig(DefaultFilterConfig())
	single := func(f string) string { return "aformat=channel_layouts=mono," + f }
	for in, opts := range inputs {
		tmp := generateTestAudio(t, opts)
		wav := filepath.Join(dir, in+".wav")
		data, _ := os.ReadFile(tmp)
		_ = os.WriteFile(wav, data, 0o644)
		cleanupTestAudio(t, tmp)
		dumpGolden(t, dir, in+"_pass1", wav, pass1.BuildFilterSpec(), false)
		dumpGolden(t, dir, in+"_pass2", wav, pass2.BuildFilterSpec(), true)
		// one filter at a time: the per-filter parity the libavfilter restatements still lack
		for name, f := range map[string]string{
			"highpass": pass2.buildRumbleHighpassFilter(), "lowpass": pass2.buildBandlimitLowPassFilter(),
			"anlmdn": "anlmdn=s=0.00001:p=0.0060:r=0.0020:m=3", "afftdn": "afftdn=nr=12:nt=w:tn=1",
			"afftdn_static": "afftdn=nr=12:nt=w:tn=0:nf=-58", "agate": pass2.buildSpeechGateFilter(),
			"acompressor": pass2.buildLevellingCompressorFilter(), "deesser": "deesser=i=0.60:m=0.50:f=0.80",
			"alimiter":  "alimiter=limit=0.501187:attack=5:release=100:level_in=1:level_out=1:level=0:latency=1:asc=1:asc_level=0.8",
			"adeclick":  pass2.buildAdeclickFilter(),
			"loudnorm":  "loudnorm=I=-16.0:TP=-1.0:LRA=20.0:dual_mono=true:print_format=json",
			"bandbody":  "atrim=start=2.000000:duration=10.000000,asetpts=PTS-STARTPTS,highpass=f=1000.000000:p=2,lowpass=f=3000.000000:p=2,astats=metadata=1:measure_perchannel=0",
			"bandtop":   "atrim=start=2.000000:duration=10.000000,asetpts=PTS-STARTPTS,highpass=f=19595.917942:p=2,lowpass=f=29393.876913:p=2,astats=metadata=1:measure_perchannel=0",
			"region":    "atrim=start=2.000000:duration=10.000000,asetpts=PTS-STARTPTS,astats=metadata=1:measure_perchannel=0,aspectralstats=measure=all,ebur128=metadata=1:peak=sample+true",
		} {
			dumpGolden(t, dir, in+"_"+name, wav, single(f), true)
		}
	}
}
