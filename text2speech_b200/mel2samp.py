"""Caller-side data helpers of the vocoder (reference: waveglow/mel2samp.py).

``files_to_list`` / ``load_wav_to_torch`` / ``MAX_WAV_VALUE`` are what ``waveglow/inference.py``
imports (mel2samp.py:40-57); ``Mel2Samp`` is the (mel, audio-segment) dataset of
``waveglow/train.py`` (mel2samp.py:60-108) with the mel computed by this package's GPU
``TacotronSTFT`` instead of a per-item CPU conv.  ``mel_batch`` is the batched form the
reference lacks: many equal-length waveforms -> one ``mel_spectrogram`` call.
"""
from __future__ import annotations

import random

import torch

from .layers import TacotronSTFT

MAX_WAV_VALUE = 32768.0                     # mel2samp.py:40


def files_to_list(filename):
    """Text file of file names -> list of file names (mel2samp.py:42-50)."""
    with open(filename, encoding="utf-8") as f:
        files = f.readlines()
    return [f.rstrip() for f in files]


def load_wav_to_torch(full_path):
    """wav file -> (float tensor of raw sample values, sampling rate)  (mel2samp.py:52-57)."""
    from scipy.io.wavfile import read
    sampling_rate, data = read(full_path)
    return torch.from_numpy(data).float(), sampling_rate


class Mel2Samp(torch.utils.data.Dataset):
    """Spectrogram / audio-segment pairs (mel2samp.py:60-108).  ``device`` is where the STFT kernels run."""

    def __init__(self, training_files, segment_length, filter_length, hop_length, win_length, sampling_rate,
                 mel_fmin, mel_fmax, device="cuda"):
        self.audio_files = files_to_list(training_files)
        random.seed(1234)
        random.shuffle(self.audio_files)
        self.stft = TacotronSTFT(filter_length=filter_length, hop_length=hop_length, win_length=win_length,
                                 sampling_rate=sampling_rate, mel_fmin=mel_fmin, mel_fmax=mel_fmax)
        self.segment_length = segment_length
        self.sampling_rate = sampling_rate
        self.device = torch.device(device)

    def get_mel(self, audio):
        """audio [N] in raw int16 units -> log-mel [80, N // hop + 1] on ``self.device``."""
        audio_norm = (audio / MAX_WAV_VALUE).unsqueeze(0).to(self.device)
        return torch.squeeze(self.stft.mel_spectrogram(audio_norm), 0)

    def mel_batch(self, audio):
        """audio [B, N] in raw int16 units -> log-mel [B, 80, N // hop + 1] in one launch sequence."""
        return self.stft.mel_spectrogram((audio / MAX_WAV_VALUE).to(self.device))

    def take_segment(self, audio):
        """Random ``segment_length`` crop, zero padded when the file is shorter (mel2samp.py:93-99)."""
        if audio.size(0) >= self.segment_length:
            start = random.randint(0, audio.size(0) - self.segment_length)
            return audio[start: start + self.segment_length]
        return torch.nn.functional.pad(audio, (0, self.segment_length - audio.size(0)), "constant").data

    def __getitem__(self, index):
        filename = self.audio_files[index]
        audio, sampling_rate = load_wav_to_torch(filename)
        if sampling_rate != self.sampling_rate:
            raise ValueError("{} SR doesn't match target {} SR".format(sampling_rate, self.sampling_rate))
        audio = self.take_segment(audio)
        mel = self.get_mel(audio)
        return (mel, audio / MAX_WAV_VALUE)

    def __len__(self):
        return len(self.audio_files)


# ===================================================================
# Takes a list of clean audio files and writes their mel spectrograms (mel2samp.py:110-142)
#   python -m text2speech_b200.mel2samp -f files.txt -c config.json -o mels/ [--batch 32]
# Same outputs as the reference (<output_dir>/<wav name>.pt holding a [80, frames] float tensor); files of equal
# length are batched through one mel_spectrogram call.
# ===================================================================
def main(filelist_path, config, output_dir, batch=32):
    import json
    import os
    with open(config) as f:
        data_config = json.loads(f.read())["data_config"]
    mel2samp = Mel2Samp(**data_config)
    filepaths = files_to_list(filelist_path)
    if not os.path.isdir(output_dir):
        os.makedirs(output_dir)
        os.chmod(output_dir, 0o775)
    audios = [load_wav_to_torch(p)[0] for p in filepaths]
    order = sorted(range(len(audios)), key=lambda i: audios[i].shape[0])
    written = []
    pos = 0
    while pos < len(order):
        group = [order[pos]]
        while (len(group) < max(1, batch) and pos + len(group) < len(order)
               and audios[order[pos + len(group)]].shape[0] == audios[group[0]].shape[0]):
            group.append(order[pos + len(group)])
        pos += len(group)
        mels = mel2samp.mel_batch(torch.stack([audios[i] for i in group])).cpu()
        for row, i in enumerate(group):
            new_filepath = output_dir + "/" + os.path.basename(filepaths[i]) + ".pt"
            torch.save(mels[row].clone(), new_filepath)
            written.append(new_filepath)
    for p in sorted(written, key=written.index):
        print(p)
    return written


if __name__ == "__main__":
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument("-f", "--filelist_path", required=True)
    parser.add_argument("-c", "--config", type=str, help="JSON file for configuration")
    parser.add_argument("-o", "--output_dir", type=str, help="Output directory")
    parser.add_argument("--batch", type=int, default=32, help="equal-length files per mel_spectrogram call")
    args = parser.parse_args()
    main(args.filelist_path, args.config, args.output_dir, args.batch)
